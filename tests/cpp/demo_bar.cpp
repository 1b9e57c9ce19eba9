// The workload of reference examples/deform_bar.cpp:30-53 (BASELINE.json configs[0]) through the C++ facade:
// bar mesh, 81 anchors pinned at rest, 81 handles rotated by pi/4 about x in one step, PrecisionType double on a
// float mesh (deform_bar.cpp:38), 10 iterations. Prints the deformed positions for comparison with the oracle.
//   usage: demo_bar <bar.obj> <constraints.txt>   (constraints: "idx x y z" per line)
#include <deform/arap.h>
#include <deform/simple_mesh.h>

#include <cstdio>
#include <fstream>

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s mesh.obj constraints.txt [iterations]\n", argv[0]); return 2; }
    typedef deform::SimpleTriMesh<float> Mesh;
    typedef deform::SimpleMeshAdapter<float> Adapter;
    typedef deform::AsRigidAsPossibleDeformation<Adapter, double> ARAP;
    Mesh mesh;
    if (!mesh.readObj(argv[1])) { std::fprintf(stderr, "Failed to read mesh\n"); return -1; }
    Adapter ma(mesh);
    ARAP arap(ma);
    std::ifstream in(argv[2]);
    int idx; double x, y, z;
    while (in >> idx >> x >> y >> z) arap.setConstraint(idx, Eigen::Vector3d(x, y, z));
    const int iterations = argc > 3 ? std::atoi(argv[3]) : 10;
    if (!arap.deform(iterations)) { std::fprintf(stderr, "deform failed: %s\n", arap.lastError()); return 1; }
    std::printf("ENERGY %.17g\n", arap.energy());
    for (int v = 0; v < ma.numberOfVertices(); ++v) {
        const Adapter::VertexType p = ma.vertexLocation(v);
        std::printf("V %.9g %.9g %.9g\n", p(0), p(1), p(2));
    }
    return 0;
}
