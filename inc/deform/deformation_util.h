// deform/deformation_util.h -- drives a group of handle vertices with one rigid transform.
//
// Public surface of the reference's deform::DeformationUtil<MeshType> (reference inc/deform/deformation_util.h:19-64):
//     DeformationUtil(mesh, handlesBegin, handlesEnd, origin = identity)   -- remembers where the handles are NOW
//     updateConstraints(t, arap)                                          -- pins handle i at  origin * t * origin^-1 * p0_i
// Internals differ: the handle positions are kept as a flat x,y,z array and every update folds origin, t and origin^-1 into one
// 3x4 matrix first (deform/detail/se3_spline.h, the arithmetic the C ABI's arap_rigid_conjugate / arap_set_rigid_constraints
// share), so an update is 12 multiply-adds per handle and the solver sees the handles as one batch of pending constraints.
#ifndef DEFORM_DEFORMATION_UTIL_H
#define DEFORM_DEFORMATION_UTIL_H

#include <deform/arap.h>
#include <deform/detail/se3_spline.h>

#include <cstddef>
#include <vector>

namespace deform {

template <class MeshType>
class DeformationUtil {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    typedef MeshType Mesh;
    typedef typename MeshType::Scalar Scalar;                               ///< the mesh's floating point type
    typedef Eigen::Transform<Scalar, 3, Eigen::Affine> Transform;           ///< rigid transforms are passed as 4x4 affine maps

    template <class HandleIterator>
    DeformationUtil(const Mesh &mesh, HandleIterator handlesBegin, HandleIterator handlesEnd, const Transform &origin = Transform::Identity()) {
        toArray(origin, _frame);
        for (HandleIterator it = handlesBegin; it != handlesEnd; ++it) {
            const int v = static_cast<int>(*it);
            const auto p = mesh.vertexLocation(v);           // the mesh concept only promises something indexable with (0..2)
            _vertex.push_back(v);
            _rest.push_back(p(0));
            _rest.push_back(p(1));
            _rest.push_back(p(2));
        }
    }

    /** One setConstraint per handle with the handle's remembered position moved by `t`, `t` being expressed in the frame `origin`. */
    template <class ARAP>
    void updateConstraints(const Transform &t, ARAP &arap) const {
        Scalar motion[16], world[16];
        toArray(t, motion);
        detail::conjugate_rigid<Scalar>(_frame, motion, world);              // origin * t * origin^-1 (isometry inverse)
        for (std::size_t k = 0; k < _vertex.size(); ++k) {
            const Scalar *p = &_rest[3 * k];
            Eigen::Matrix<Scalar, 3, 1> target;
            for (int r = 0; r < 3; ++r) target(r) = world[4 * r] * p[0] + world[4 * r + 1] * p[1] + world[4 * r + 2] * p[2] + world[4 * r + 3];
            arap.setConstraint(_vertex[k], target);
        }
    }

    /** The handle vertex ids, in the order they were given. */
    const std::vector<int> &handles() const { return _vertex; }

private:
    static void toArray(const Transform &T, Scalar out[16]) {
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) out[4 * r + c] = T.matrix()(r, c);
    }

    Scalar _frame[16];                 // `origin`, row-major
    std::vector<int> _vertex;          // handle ids
    std::vector<Scalar> _rest;         // their positions at construction, x,y,z each
};

}  // namespace deform

#endif
