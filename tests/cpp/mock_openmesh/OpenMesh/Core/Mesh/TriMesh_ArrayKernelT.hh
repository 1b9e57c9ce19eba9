// TEST DOUBLE of the part of OpenMesh's public API that deform/openmesh_adapter.h and the reference's demos touch
// (reference inc/deform/openmesh_adapter.h:18-19,55,66-113; examples/deform_*.cpp: IO::read_mesh, point, set_point,
// vertex_handle, face_handle, fv_begin / cfv_iter, n_vertices, n_faces, add_vertex, add_face). OpenMesh itself is not
// installed in the build image; this header exists ONLY so that the adapter is compiled and exercised by the test-suite
// (tests/cpp/test_openmesh_adapter.cpp). It is written from OpenMesh's documented interface, not from its sources, and
// is NOT part of the product: a real build puts OpenMesh's include directory on the path instead of tests/cpp/mock_openmesh.
// Like the real array kernel it keeps the points in ONE contiguous array of VectorT<Scalar,3> in vertex-index order.
#ifndef MOCK_OPENMESH_TRIMESH_ARRAYKERNELT_HH
#define MOCK_OPENMESH_TRIMESH_ARRAYKERNELT_HH

#include <cstddef>
#include <vector>

namespace OpenMesh {

template <class S, int N>
class VectorT {
public:
    typedef S value_type;
    VectorT() { for (int i = 0; i < N; ++i) _v[i] = S(0); }
    VectorT(S x, S y, S z) { static_assert(N == 3, "3-vector constructor"); _v[0] = x; _v[1] = y; _v[2] = z; }
    S &operator[](size_t i) { return _v[i]; }
    const S &operator[](size_t i) const { return _v[i]; }
    S *data() { return _v; }
    const S *data() const { return _v; }
private:
    S _v[N];
};
typedef VectorT<float, 3> Vec3f;
typedef VectorT<double, 3> Vec3d;

struct DefaultTraits { typedef Vec3f Point; };

class BaseHandle {
public:
    explicit BaseHandle(int idx = -1) : _idx(idx) {}
    int idx() const { return _idx; }
    bool is_valid() const { return _idx >= 0; }
private:
    int _idx;
};
struct VertexHandle : public BaseHandle { explicit VertexHandle(int idx = -1) : BaseHandle(idx) {} };
struct FaceHandle : public BaseHandle { explicit FaceHandle(int idx = -1) : BaseHandle(idx) {} };

template <class Traits = DefaultTraits>
class TriMesh_ArrayKernelT {
public:
    typedef typename Traits::Point Point;
    typedef typename Point::value_type Scalar;
    typedef OpenMesh::VertexHandle VertexHandle;
    typedef OpenMesh::FaceHandle FaceHandle;

    /** Circulator over the three vertices of a face (FaceVertexIter / ConstFaceVertexIter). */
    class ConstFaceVertexIter {
    public:
        ConstFaceVertexIter(const int *ids, int pos) : _ids(ids), _pos(pos), _cur(pos < 3 ? ids[pos] : -1) {}
        bool is_valid() const { return _pos < 3; }
        ConstFaceVertexIter &operator++() { ++_pos; _cur = VertexHandle(_pos < 3 ? _ids[_pos] : -1); return *this; }
        const VertexHandle &operator*() const { return _cur; }
        const VertexHandle *operator->() const { return &_cur; }
    private:
        const int *_ids;
        int _pos;
        VertexHandle _cur;
    };
    typedef ConstFaceVertexIter FaceVertexIter;

    VertexHandle vertex_handle(unsigned int i) const { return VertexHandle(i < _points.size() ? (int)i : -1); }
    FaceHandle face_handle(unsigned int i) const { return FaceHandle(i < _faces.size() / 3 ? (int)i : -1); }
    const Point &point(VertexHandle vh) const { return _points[(size_t)vh.idx()]; }
    Point &point(VertexHandle vh) { return _points[(size_t)vh.idx()]; }
    void set_point(VertexHandle vh, const Point &p) { _points[(size_t)vh.idx()] = p; }
    size_t n_vertices() const { return _points.size(); }
    size_t n_faces() const { return _faces.size() / 3; }
    VertexHandle add_vertex(const Point &p) { _points.push_back(p); return VertexHandle((int)_points.size() - 1); }
    FaceHandle add_face(VertexHandle a, VertexHandle b, VertexHandle c) {
        _faces.push_back(a.idx()); _faces.push_back(b.idx()); _faces.push_back(c.idx());
        return FaceHandle((int)_faces.size() / 3 - 1);
    }
    ConstFaceVertexIter cfv_iter(FaceHandle fh) const { return ConstFaceVertexIter(&_faces[3 * (size_t)fh.idx()], 0); }
    ConstFaceVertexIter cfv_begin(FaceHandle fh) const { return cfv_iter(fh); }
    ConstFaceVertexIter fv_begin(FaceHandle fh) const { return cfv_iter(fh); }
    void clear() { _points.clear(); _faces.clear(); }

private:
    std::vector<Point> _points;
    std::vector<int> _faces;
};

}  // namespace OpenMesh

#endif
