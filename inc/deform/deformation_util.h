// deform/deformation_util.h -- moves a set of handle vertices rigidly and feeds them to the solver.
// Same interface as the reference's deform::DeformationUtil<MeshType> (reference inc/deform/deformation_util.h:19-64).
#ifndef DEFORM_DEFORMATION_UTIL_H
#define DEFORM_DEFORMATION_UTIL_H

#include <deform/arap.h>
#include <deform/trajectory.h>

#include <vector>

namespace deform {

template <class MeshType>
class DeformationUtil {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    typedef MeshType Mesh;
    /** Floating point precision of the mesh. */
    typedef typename MeshType::Scalar Scalar;
    /** Transformation matrix type. */
    typedef Eigen::Transform<Scalar, 3, Eigen::Affine> Transform;

    /** Remembers where the handles are NOW; `origin` is the frame the later transforms are expressed in. */
    template <class HandleIterator>
    DeformationUtil(const Mesh &mesh, HandleIterator handlesBegin, HandleIterator handlesEnd, const Transform &origin = Transform::Identity())
        : _handles(handlesBegin, handlesEnd), _origin(origin), _originInv(origin.inverse(Eigen::Isometry)) {
        _points.reserve(_handles.size());
        for (size_t i = 0; i < _handles.size(); ++i) _points.push_back(mesh.vertexLocation(_handles[i]));
    }

    /** setConstraint(handle_i, origin * t * origin^-1 * p_i) for every handle (reference deformation_util.h:48-57). */
    template <class ARAP>
    void updateConstraints(const Transform &t, ARAP &arap) {
        const Transform tabs = _origin * t * _originInv;
        for (size_t i = 0; i < _handles.size(); ++i) arap.setConstraint(_handles[i], Eigen::Matrix<Scalar, 3, 1>(tabs * _points[i]));
    }

    const std::vector<int> &handles() const { return _handles; }

private:
    std::vector<int> _handles;
    std::vector<Eigen::Matrix<Scalar, 3, 1> > _points;
    Transform _origin, _originInv;
};

}  // namespace deform

#endif
