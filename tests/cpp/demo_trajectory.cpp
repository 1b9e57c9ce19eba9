// The call sequence of reference examples/deform_trajectory.cpp:46-86 through the C++ facade, without the viewer:
// bar mesh (float), ARAP in double, anchors pinned at rest, the handles driven by DeformationUtil along a TrajectorySE3<float>
// through the reference's four key poses; every frame = updateConstraints(trajectory(t)) + deform(n) (the dirty protocol re-reads
// the deformed mesh as the new rest pose each frame, arap.h:84,102-120). Prints the mesh after every frame.
//   usage: demo_trajectory <bar.obj> <anchors.txt> <handles.txt> <iterations per frame> <t1> [t2 ...]
#include <deform/arap.h>
#include <deform/deformation_util.h>
#include <deform/simple_mesh.h>
#include <deform/trajectory.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static std::vector<int> readInts(const char *path) {
    std::vector<int> v;
    std::ifstream in(path);
    int x;
    while (in >> x) v.push_back(x);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s mesh.obj anchors.txt handles.txt iterations t1 [t2 ...]\n", argv[0]); return 2; }
    typedef deform::SimpleTriMesh<float> Mesh;
    typedef deform::SimpleMeshAdapter<float> Adapter;
    typedef deform::AsRigidAsPossibleDeformation<Adapter, double> ARAP;
    typedef deform::TrajectorySE3<float> Trajectory;
    typedef deform::DeformationUtil<Adapter> DeformationUtil;

    Mesh mesh;
    if (!mesh.readObj(argv[1])) { std::fprintf(stderr, "Failed to read mesh\n"); return -1; }
    const std::vector<int> anchors = readInts(argv[2]), handles = readInts(argv[3]);
    const int iterations = std::atoi(argv[4]);

    Adapter ma(mesh);
    ARAP arap(ma);
    for (size_t i = 0; i < anchors.size(); ++i) arap.setConstraint(anchors[i], ma.vertexLocation(anchors[i]));   // deform_trajectory.cpp:56-59

    Trajectory trajectory;                                                                                       // :62-69
    Trajectory::Transform prev;
    prev = trajectory.addKeyPose(Trajectory::Transform::Identity());
    prev = trajectory.addKeyPose(prev * Eigen::Translation3f(1.f, 0.f, 0.f));
    prev = trajectory.addKeyPose(prev * Eigen::Translation3f(2.f, 0.f, 0.f));
    prev = trajectory.addKeyPose(prev * Eigen::AngleAxisf((float)M_PI / 2.f, Eigen::Vector3f::UnitX()));

    DeformationUtil dutil(ma, handles.begin(), handles.end(), trajectory(0.0));                                  // :71

    for (int a = 5; a < argc; ++a) {
        const float t = (float)std::atof(argv[a]);
        dutil.updateConstraints(trajectory(t), arap);                                                            // :84
        if (!arap.deform(iterations)) { std::fprintf(stderr, "deform failed: %s\n", arap.lastError()); return 1; } // :86
        std::printf("FRAME %.9g ENERGY %.17g\n", t, arap.energy());
        for (int v = 0; v < ma.numberOfVertices(); ++v) {
            const Adapter::VertexType p = ma.vertexLocation(v);
            std::printf("V %.9g %.9g %.9g\n", p(0), p(1), p(2));
        }
    }
    return 0;
}
