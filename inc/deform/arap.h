// deform/arap.h -- As-rigid-as-possible surface deformation, B200 edition.
//
// Same public surface as cheind/mesh-deform's header of the same name (reference inc/deform/arap.h:49-138):
//
//     deform::AsRigidAsPossibleDeformation<MeshType, PrecisionType = MeshType::Scalar>
//         typedef Scalar, Mesh, Index
//         AsRigidAsPossibleDeformation(Mesh &mesh)                       // mesh must outlive the object
//         template<class S> void setConstraint(Index, const Eigen::Matrix<S,3,1> &)
//         bool deform(Index numberOfIterations)
//         template<class> friend class PrivateAccessor                   // tests/accessor.h still works
//
// so existing callers (the demos, DeformationUtil, the reference's own tests) compile unchanged. What
// changed is everything behind it: this class is a thin host wrapper that forwards to the C ABI of
// libarap_b200.so (include/arap_b200.h), where cotan weights + CSR, the local step, the right-hand
// side and the global solve run as CUDA kernels on the GPU. There is no CPU implementation in here.
//
// Semantics kept from the reference: the `_dirty` protocol -- any setConstraint makes
// the next deform() re-read the rest pose FROM THE MESH, rebuild weights, reset rotations and set the system
// up again (arap.h:102-120); deform() returns true early without touching the mesh when nothing is
// constrained (arap.h:113-114); deform(0) snaps constrained vertices to their targets; consecutive deform()
// calls continue from the previous state; constraints accumulate; PrecisionType may differ from the mesh scalar.
#ifndef DEFORM_ARAP_H
#define DEFORM_ARAP_H

#include <arap_b200.h>
#include <deform/detail/linalg.h>

#include <vector>

namespace deform {

namespace detail {
// Optional extension of the mesh concept (reference inc/deform/openmesh_adapter.h:49-118 has five members, called once
// per vertex / face): a mesh whose vertex positions (and faces) are contiguous x,y,z (v0,v1,v2) arrays may also offer
//     const Scalar *vertexData() const;   Scalar *vertexData();   const int *faceData() const;
// and the solver then hands those pointers straight to the engine instead of making V (or F) accessor calls per
// deform(). Detected at compile time; meshes with only the five required members work unchanged.
template <class M> struct has_vertex_data {
    template <class T> static char test(decltype(static_cast<const typename T::Scalar *>(static_cast<const T *>(nullptr)->vertexData())) *,
                                        decltype(static_cast<typename T::Scalar *>(static_cast<T *>(nullptr)->vertexData())) *);
    template <class T> static long test(...);
    static const bool value = sizeof(test<M>(nullptr, nullptr)) == sizeof(char);
};
template <class M> struct has_face_data {
    template <class T> static char test(decltype(static_cast<const int *>(static_cast<const T *>(nullptr)->faceData())) *);
    template <class T> static long test(...);
    static const bool value = sizeof(test<M>(nullptr)) == sizeof(char);
};
template <bool B> struct bool_tag {};
}  // namespace detail

/** Forward declaration of a class that may look into AsRigidAsPossibleDeformation (used by tests). */
template <class T> class PrivateAccessor;

template <class MeshType, class PrecisionType = typename MeshType::Scalar>
class AsRigidAsPossibleDeformation {
public:
    /** Floating point precision used in calculations (float or double). */
    typedef PrecisionType Scalar;
    /** Mesh type we are working on. */
    typedef MeshType Mesh;
    /** Type for indices. */
    typedef int Index;

    /** Construct from mesh; reads the topology once (reference arap.h:66-70,149-155). */
    explicit AsRigidAsPossibleDeformation(Mesh &mesh) : _mesh(mesh), _handle(nullptr), _dirty(true), _edgeWeights(this) {
        static_assert(sizeof(Scalar) == 4 || sizeof(Scalar) == 8, "PrecisionType must be float or double");
        createHandle(detail::bool_tag<detail::has_face_data<Mesh>::value && sizeof(int) == sizeof(int32_t)>());
    }

    ~AsRigidAsPossibleDeformation() { arap_destroy(_handle); }

    AsRigidAsPossibleDeformation(const AsRigidAsPossibleDeformation &) = delete;
    AsRigidAsPossibleDeformation &operator=(const AsRigidAsPossibleDeformation &) = delete;

    /** Pin vertex `vidx` at `loc` (anchors and handles are the same thing). Reference arap.h:81-85. */
    template <class S, class Vec3>
    void setConstraintXYZ(Index vidx, const Vec3 &loc) {
        _pendingIdx.push_back(vidx);
        _pendingLoc.push_back((Scalar)loc(0));
        _pendingLoc.push_back((Scalar)loc(1));
        _pendingLoc.push_back((Scalar)loc(2));
        _dirty = true;
    }
    template <class S>
    void setConstraint(Index vidx, const Eigen::Matrix<S, 3, 1> &loc) { setConstraintXYZ<S>(vidx, loc); }

    /** Run `numberOfIterations` ARAP iterations; see the reference's contract at arap.h:87-101. */
    bool deform(Index numberOfIterations) {
        if (!_handle) return false;
        typedef typename Mesh::Scalar MeshScalar;
        const detail::bool_tag<detail::has_vertex_data<Mesh>::value> bulk;
        MeshScalar *xyz = vertexBuffer(bulk);               // the mesh's own storage when it offers vertexData()

        if (_dirty) {
            if (!_pendingIdx.empty()) {
                // one call for everything queued since the last deform(); a vertex named twice keeps its LAST location (arap.h:83)
                const int rc = arap_set_constraints(_handle, (int32_t)_pendingIdx.size(), _pendingIdx.data(), _pendingLoc.data(), (int32_t)sizeof(Scalar));
                _pendingIdx.clear();                         // also on failure: a bad index must not poison every later deform()
                _pendingLoc.clear();
                if (rc != ARAP_OK) return false;
            }
            readVertices(xyz, bulk);                         // initializeMeshGeometry (arap.h:162-168)
            const int rc = arap_prepare(_handle, xyz, (int32_t)sizeof(MeshScalar));
            if (rc == ARAP_UNCONSTRAINED) return true;       // arap.h:113-114 (stays dirty, no write-back)
            if (rc != ARAP_OK) return false;                 // arap.h:116-117
            _dirty = false;
        }

        const int it = arap_iterate(_handle, numberOfIterations);
        if (it < 0) return false;

        if (arap_get_positions(_handle, xyz, (int32_t)sizeof(MeshScalar)) != ARAP_OK) return false;
        writeVertices(xyz, bulk);                            // write-back (arap.h:133-135)
        return it == ARAP_OK;                                // ARAP_NOT_CONVERGED: unusable system, the reference's `false` (arap.h:116-117)
    }

    /** ARAP energy of the current state (not part of the reference API; Sorkine & Alexa 2007, eq. 3). */
    double energy() const {
        double e = 0;
        return (_handle && arap_energy(_handle, &e) == ARAP_OK) ? e : 0.0;
    }

    /** Message of the last engine error (not part of the reference API). */
    const char *lastError() const { return _handle ? arap_last_error(_handle) : arap_create_error(); }

    /** The engine handle, for the extended C API (solver statistics, profiling). */
    arap_handle *handle() const { return _handle; }

private:
    template <class> friend class PrivateAccessor;

    // ---- mesh ingest / write-back: bulk pointers when the mesh offers them, the five-member concept otherwise ----
    void createHandle(detail::bool_tag<true>) {
        if (arap_create(reinterpret_cast<const int32_t *>(_mesh.faceData()), _mesh.numberOfFaces(), _mesh.numberOfVertices(),
                        (int32_t)sizeof(Scalar), nullptr, &_handle) != ARAP_OK)
            _handle = nullptr;                               // deform() then returns false; lastError() has arap_create_error()
    }
    void createHandle(detail::bool_tag<false>) {             // initializeMeshTopology (arap.h:149-155)
        const Index nF = _mesh.numberOfFaces();
        std::vector<int32_t> faces(3 * (size_t)nF);
        for (Index f = 0; f < nF; ++f) {
            const auto vids = _mesh.face(f);
            faces[3 * (size_t)f + 0] = vids(0);
            faces[3 * (size_t)f + 1] = vids(1);
            faces[3 * (size_t)f + 2] = vids(2);
        }
        if (arap_create(faces.data(), nF, _mesh.numberOfVertices(), (int32_t)sizeof(Scalar), nullptr, &_handle) != ARAP_OK)
            _handle = nullptr;
    }
    typename Mesh::Scalar *vertexBuffer(detail::bool_tag<true>) { return _mesh.vertexData(); }
    typename Mesh::Scalar *vertexBuffer(detail::bool_tag<false>) {
        _buffer.resize(3 * (size_t)_mesh.numberOfVertices() * sizeof(typename Mesh::Scalar));
        return reinterpret_cast<typename Mesh::Scalar *>(_buffer.data());
    }
    void readVertices(typename Mesh::Scalar *, detail::bool_tag<true>) {}
    void readVertices(typename Mesh::Scalar *xyz, detail::bool_tag<false>) {
        const Index nV = _mesh.numberOfVertices();
        for (Index v = 0; v < nV; ++v) {
            const auto p = _mesh.vertexLocation(v);
            xyz[3 * (size_t)v + 0] = p(0);
            xyz[3 * (size_t)v + 1] = p(1);
            xyz[3 * (size_t)v + 2] = p(2);
        }
    }
    void writeVertices(const typename Mesh::Scalar *, detail::bool_tag<true>) {}
    void writeVertices(const typename Mesh::Scalar *xyz, detail::bool_tag<false>) {
        typedef typename Mesh::Scalar MeshScalar;
        const Index nV = _mesh.numberOfVertices();
        for (Index i = 0; i < nV; ++i)
            _mesh.vertexLocation(i, Eigen::Matrix<MeshScalar, 3, 1>(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]));
    }

#ifdef DEFORM_HAVE_EIGEN
    typedef Eigen::SparseMatrix<Scalar, Eigen::RowMajor> SparseMatrix;
#else
    typedef detail::CsrMatrix<Scalar> SparseMatrix;
#endif

    // `_edgeWeights` lives on the device. This member keeps the reference's name and converts to the
    // reference's type on demand, so `return a._edgeWeights;` in tests/accessor.h:18-20 keeps compiling.
    class DeviceEdgeWeights {
    public:
        explicit DeviceEdgeWeights(const AsRigidAsPossibleDeformation *owner) : _owner(owner) {}
        operator SparseMatrix() const {
            const int n = _owner->_mesh.numberOfVertices();
            int32_t nnz = 0;
            if (!_owner->_handle || arap_get_csr_nnz(_owner->_handle, &nnz) != ARAP_OK)
                return SparseMatrix(detail::CsrMatrix<Scalar>(n, n, std::vector<int>((size_t)n + 1, 0), std::vector<int>(), std::vector<Scalar>()));
            std::vector<int> rowptr((size_t)n + 1), colidx((size_t)nnz);
            std::vector<Scalar> values((size_t)nnz);
            arap_get_csr(_owner->_handle, rowptr.data(), colidx.data(), values.data());
            return SparseMatrix(detail::CsrMatrix<Scalar>(n, n, std::move(rowptr), std::move(colidx), std::move(values)));
        }
    private:
        const AsRigidAsPossibleDeformation *_owner;
    };

    Mesh &_mesh;
    arap_handle *_handle;
    bool _dirty;
    DeviceEdgeWeights _edgeWeights;
    std::vector<int32_t> _pendingIdx;
    std::vector<Scalar> _pendingLoc;
    std::vector<unsigned char> _buffer;
};

}  // namespace deform

#endif
