// Restatement of reference tests/test_trajectory.cpp:17-39 against deform/trajectory.h (host only, no GPU),
// followed by a dump of interior samples that the Python test compares with the oracle's golden samples.
#include "mini_test.h"

#include <deform/trajectory.h>
#include <deform/deformation_util.h>
#include <cstdio>
#include <iostream>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

TEST_CASE("trajectory")
{
    typedef deform::TrajectorySE3<float> Trajectory;

    Trajectory path;

    // time 0
    Trajectory::Transform pose0 = Trajectory::Transform::Identity();
    path.addKeyPose(pose0);

    Trajectory::Transform pose1 = Eigen::Translation3f(0, 0, 1) * pose0;
    path.addKeyPose(pose1);

    Trajectory::Transform pose2 = Eigen::Translation3f(0, 0, 1) * Eigen::AngleAxisf((float)M_PI / 4.f, Eigen::Vector3f::UnitX()) * pose1;
    path.addKeyPose(pose2);

    Trajectory::Transform pose3 = Eigen::Translation3f(0, 0, 1) * pose2;
    path.addKeyPose(pose3);

    REQUIRE(path(0.f).matrix().isApprox(pose0.matrix(), 1e-3f));
    REQUIRE(path(1.f).matrix().isApprox(pose3.matrix(), 1e-3f));
}

// Same key poses in double; prints 33 samples (u, 16 matrix entries) for comparison with tests/golden/trajectory_golden.npz.
TEST_CASE("trajectory_samples_double")
{
    typedef deform::TrajectorySE3<double> Trajectory;
    Trajectory path;
    Trajectory::Transform pose0 = Trajectory::Transform::Identity();
    Trajectory::Transform pose1 = Eigen::Translation3d(0, 0, 1) * pose0;
    Trajectory::Transform pose2 = Eigen::Translation3d(0, 0, 1) * Eigen::AngleAxisd(M_PI / 4., Eigen::Vector3d::UnitX()) * pose1;
    Trajectory::Transform pose3 = Eigen::Translation3d(0, 0, 1) * pose2;
    path.addKeyPose(pose0); path.addKeyPose(pose1); path.addKeyPose(pose2); path.addKeyPose(pose3);
    for (int k = 0; k <= 32; ++k) {
        const double u = k / 32.0;
        Trajectory::Transform T = path(u);
        std::printf("SAMPLE %.17g", u);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) std::printf(" %.17g", T.matrix()(i, j));
        std::printf("\n");
    }
    std::vector<double> batch = path.sample(5);
    REQUIRE(batch.size() == 80);
    Trajectory::Transform mid = path(0.5);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) REQUIRE(std::fabs(batch[32 + 4 * i + j] - mid.matrix()(i, j)) < 1e-14);
}

// DeformationUtil arithmetic (reference deformation_util.h:48-57) with a recording stand-in for the solver.
struct RecordingSolver {
    std::vector<int> idx;
    std::vector<Eigen::Vector3d> loc;
    template <class S> void setConstraint(int i, const Eigen::Matrix<S, 3, 1> &p) { idx.push_back(i); loc.push_back(Eigen::Vector3d(p(0), p(1), p(2))); }
};
struct TinyMesh {
    typedef double Scalar;
    Eigen::Vector3d vertexLocation(int i) const { return Eigen::Vector3d(1.0 + i, 2.0, -0.5 * i); }
};

TEST_CASE("deformation_util")
{
    typedef deform::DeformationUtil<TinyMesh> Util;
    TinyMesh mesh;
    const int handles[3] = {0, 2, 5};
    Util::Transform origin = Eigen::Translation3d(1, 0, 0) * Util::Transform(Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitZ()));
    Util util(mesh, handles, handles + 3, origin);
    Util::Transform t = Eigen::Translation3d(0, 0.5, 0) * Util::Transform(Eigen::AngleAxisd(0.7, Eigen::Vector3d::UnitX()));
    RecordingSolver rec;
    util.updateConstraints(t, rec);
    REQUIRE(rec.idx.size() == 3);
    Util::Transform tabs = origin * t * origin.inverse(Eigen::Isometry);
    for (int k = 0; k < 3; ++k) {
        REQUIRE(rec.idx[k] == handles[k]);
        REQUIRE((rec.loc[k] - tabs * mesh.vertexLocation(handles[k])).norm() < 1e-14);
        std::printf("HANDLE %d %.17g %.17g %.17g\n", handles[k], rec.loc[k](0), rec.loc[k](1), rec.loc[k](2));
    }
    for (int i = 0; i < 4; ++i) { std::printf("ORIGIN"); for (int j = 0; j < 4; ++j) std::printf(" %.17g", origin.matrix()(i, j)); std::printf("\n"); }
    for (int i = 0; i < 4; ++i) { std::printf("T"); for (int j = 0; j < 4; ++j) std::printf(" %.17g", t.matrix()(i, j)); std::printf("\n"); }
}

MINI_TEST_MAIN
