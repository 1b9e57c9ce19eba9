// The workload of reference examples/deform_bar.cpp:30-53 (BASELINE.json configs[0]) through the C++ facade:
// bar mesh, 81 anchors pinned at rest, 81 handles rotated by pi/4 about x in one step, PrecisionType double on a
// float mesh (deform_bar.cpp:38), 10 iterations. Prints the deformed positions for comparison with the oracle.
//   usage: demo_bar <bar.obj> <constraints.txt> [iterations] [minimal]   (constraints: "idx x y z" per line)
// With "minimal" the mesh adapter offers only the five members of the reference's mesh concept
// (openmesh_adapter.h:55,74-113), so the solver takes its per-vertex accessor path instead of the bulk pointers.
#include <deform/arap.h>
#include <deform/simple_mesh.h>

#include <cstdio>
#include <cstring>
#include <fstream>

// exactly the reference's mesh concept, nothing more
class FiveMemberAdapter {
public:
    typedef float Scalar;
    typedef Eigen::Matrix<float, 3, 1> VertexType;
    typedef Eigen::Matrix<int, 3, 1> FaceType;
    explicit FiveMemberAdapter(deform::SimpleTriMesh<float> &m) : _inner(m) {}
    VertexType vertexLocation(int i) const { return _inner.vertexLocation(i); }
    void vertexLocation(int i, const VertexType &v) { _inner.vertexLocation(i, v); }
    FaceType face(int f) const { return _inner.face(f); }
    int numberOfFaces() const { return _inner.numberOfFaces(); }
    int numberOfVertices() const { return _inner.numberOfVertices(); }
private:
    deform::SimpleMeshAdapter<float> _inner;
};
static_assert(deform::detail::has_vertex_data<deform::SimpleMeshAdapter<float> >::value, "SimpleMeshAdapter offers the bulk accessors");
static_assert(deform::detail::has_face_data<deform::SimpleMeshAdapter<float> >::value, "SimpleMeshAdapter offers faceData");
static_assert(!deform::detail::has_vertex_data<FiveMemberAdapter>::value, "the five-member concept has no bulk accessors");
static_assert(!deform::detail::has_face_data<FiveMemberAdapter>::value, "the five-member concept has no faceData");

template <class Adapter>
int run(deform::SimpleTriMesh<float> &mesh, int argc, char **argv) {
    typedef deform::AsRigidAsPossibleDeformation<Adapter, double> ARAP;
    Adapter ma(mesh);
    ARAP arap(ma);
    std::ifstream in(argv[2]);
    int idx; double x, y, z;
    while (in >> idx >> x >> y >> z) arap.setConstraint(idx, Eigen::Vector3d(x, y, z));
    const int iterations = argc > 3 ? std::atoi(argv[3]) : 10;
    if (!arap.deform(iterations)) { std::fprintf(stderr, "deform failed: %s\n", arap.lastError()); return 1; }
    std::printf("ENERGY %.17g\n", arap.energy());
    for (int v = 0; v < ma.numberOfVertices(); ++v) {
        const typename Adapter::VertexType p = ma.vertexLocation(v);
        std::printf("V %.9g %.9g %.9g\n", p(0), p(1), p(2));
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s mesh.obj constraints.txt [iterations] [minimal]\n", argv[0]); return 2; }
    deform::SimpleTriMesh<float> mesh;
    if (!mesh.readObj(argv[1])) { std::fprintf(stderr, "Failed to read mesh\n"); return -1; }
    if (argc > 4 && std::strcmp(argv[4], "minimal") == 0) return run<FiveMemberAdapter>(mesh, argc, argv);
    return run<deform::SimpleMeshAdapter<float> >(mesh, argc, argv);
}
