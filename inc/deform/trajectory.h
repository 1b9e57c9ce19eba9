// deform/trajectory.h -- smooth SE(3) trajectories through key poses.
//
// Same interface as the reference's deform::TrajectorySE3<PrecisionType> (reference inc/deform/trajectory.h:33-83):
// addKeyPose(Transform) -> Transform, operator()(time in [0,1]) -> Transform. The arithmetic (SE(3) log/exp, the
// interpolating cubic B-spline) lives in deform/detail/se3_spline.h, shared with the C ABI's arap_trajectory_*.
// Host code: a trajectory is a handful of poses; it only generates constraint targets for the solver.
// `sample(n)` evaluates n poses at once for batched deformations (one pose per batch member).
#ifndef DEFORM_TRAJECTORY_H
#define DEFORM_TRAJECTORY_H

#include <deform/detail/linalg.h>
#include <deform/detail/se3_spline.h>

#include <vector>

namespace deform {

template <class PrecisionType>
class TrajectorySE3 {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    /** Floating point precision used in calculations. */
    typedef PrecisionType Scalar;
    /** Transformation matrix type. */
    typedef Eigen::Transform<Scalar, 3, Eigen::Affine> Transform;

    TrajectorySE3() {}

    /** Append a key pose; returns it (reference trajectory.h:55-59). Needs >= 4 poses before evaluation. */
    Transform addKeyPose(const Transform &transform) {
        Scalar m[16];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[4 * i + j] = transform.matrix()(i, j);
        _spline.addKeyPose(m);
        return transform;
    }

    /** Pose at `time` in [0,1] (reference trajectory.h:61-73). */
    Transform operator()(Scalar time) {
        Scalar T[16];
        _spline.pose(time, T);
        Transform out;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out.matrix()(i, j) = T[4 * i + j];
        return out;
    }

    /** 4x4 row-major matrices of n poses at times 0, 1/(n-1), ..., 1 -- one per member of a batch of deformations. */
    std::vector<Scalar> sample(int n) {
        std::vector<Scalar> out(16 * (size_t)n);
        for (int k = 0; k < n; ++k) _spline.pose(n > 1 ? Scalar(k) / Scalar(n - 1) : Scalar(0), &out[16 * (size_t)k]);
        return out;
    }

    int numberOfKeyPoses() const { return _spline.numberOfKeyPoses(); }

private:
    detail::SplineSE3<Scalar> _spline;
};

}  // namespace deform

#endif
