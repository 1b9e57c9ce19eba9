// deform/openmesh_adapter.h -- connects OpenMesh triangle meshes to the deform algorithms.
// Same names as the reference (reference inc/deform/openmesh_adapter.h:24-118): deform::convert::toEigen /
// toOpenMesh and deform::OpenMeshAdapter<Traits> with the five required members of the mesh concept.
// Needs OpenMesh and Eigen; where they are not installed (this build image) use deform/simple_mesh.h.
#ifndef DEFORM_OPENMESH_ADAPTER_H
#define DEFORM_OPENMESH_ADAPTER_H

#if defined(__has_include)
#if !__has_include(<OpenMesh/Core/Mesh/TriMesh_ArrayKernelT.hh>)
#error "deform/openmesh_adapter.h needs OpenMesh; use deform/simple_mesh.h (SimpleMeshAdapter) when OpenMesh is not installed"
#endif
#endif

#ifndef _USE_MATH_DEFINES
#define _USE_MATH_DEFINES
#endif

#include <OpenMesh/Core/IO/MeshIO.hh>
#include <OpenMesh/Core/Mesh/TriMesh_ArrayKernelT.hh>
#include <deform/detail/linalg.h>

namespace deform {

namespace convert {

template <class Scalar>
Eigen::Matrix<Scalar, 3, 1> toEigen(const OpenMesh::VectorT<Scalar, 3> &v) { return Eigen::Matrix<Scalar, 3, 1>(v[0], v[1], v[2]); }

template <class Scalar>
OpenMesh::VectorT<Scalar, 3> toOpenMesh(const Eigen::Matrix<Scalar, 3, 1> &v) { return OpenMesh::VectorT<Scalar, 3>(v(0), v(1), v(2)); }

}  // namespace convert

template <class Traits = OpenMesh::DefaultTraits>
class OpenMeshAdapter {
public:
    typedef OpenMesh::TriMesh_ArrayKernelT<Traits> Mesh;
    typedef typename Mesh::Scalar Scalar;                 // required by the solver
    typedef Eigen::Matrix<Scalar, 3, 1> VertexType;
    typedef Eigen::Matrix<int, 3, 1> FaceType;

    /** The mesh is referenced, not copied; it has to outlive the adapter. */
    explicit OpenMeshAdapter(Mesh &mesh) : _mesh(mesh) {}

    VertexType vertexLocation(int idx) const { return convert::toEigen(_mesh.point(_mesh.vertex_handle((unsigned)idx))); }
    void vertexLocation(int idx, const VertexType &v) { _mesh.set_point(_mesh.vertex_handle((unsigned)idx), convert::toOpenMesh(v)); }

    /** The three vertex ids of a face, in the mesh's own winding. */
    FaceType face(int idx) const {
        FaceType ids;
        int k = 0;
        for (auto it = _mesh.cfv_iter(_mesh.face_handle((unsigned)idx)); it.is_valid() && k < 3; ++it, ++k) ids(k) = it->idx();
        return ids;
    }
    int numberOfFaces() const { return (int)_mesh.n_faces(); }
    int numberOfVertices() const { return (int)_mesh.n_vertices(); }

    /** Optional bulk access (see deform/arap.h): OpenMesh keeps its points in one contiguous property array of
     *  VectorT<Scalar,3> in vertex-index order, so the solver can read and write all positions with one copy each
     *  instead of n_vertices() accessor calls. Faces are not stored contiguously; `face()` stays the only way. */
    const Scalar *vertexData() const {
        static_assert(sizeof(typename Mesh::Point) == 3 * sizeof(Scalar), "OpenMesh points are expected to be packed x,y,z");
        return _mesh.n_vertices() ? &_mesh.point(_mesh.vertex_handle(0u))[0] : nullptr;
    }
    Scalar *vertexData() {
        return _mesh.n_vertices() ? &_mesh.point(_mesh.vertex_handle(0u))[0] : nullptr;
    }

private:
    Mesh &_mesh;
};

}  // namespace deform

#endif
