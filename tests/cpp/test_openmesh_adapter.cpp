// deform/openmesh_adapter.h compiled and exercised: the reference's tests/test_cotan.cpp:20-54 VERBATIM in its use of
// the OpenMesh API (add_vertex / add_face / OpenMeshAdapter<>), plus the façade's error handling. OpenMesh is not in the
// build image, so the OpenMesh headers come from tests/cpp/mock_openmesh (an API-compatible test double, see the
// header there); with a real OpenMesh on the include path this file compiles unchanged. Needs a GPU to run.
#include "mini_test.h"

#include <deform/arap.h>
#include <deform/openmesh_adapter.h>
#include <deform/deformation_util.h>
#include <iostream>

#include "accessor.h"

TEST_CASE("cotan_weights_openmesh")          // reference tests/test_cotan.cpp:20-54
{
    typedef OpenMesh::TriMesh_ArrayKernelT<> Mesh;

    Mesh m;

    Mesh::VertexHandle vhandle[4];
    vhandle[0] = m.add_vertex(Mesh::Point(0.f, 0.f, 0.f));
    vhandle[1] = m.add_vertex(Mesh::Point(1.f, 0.f, 0.f));
    vhandle[2] = m.add_vertex(Mesh::Point(1.f, 1.f, 0.f));
    vhandle[3] = m.add_vertex(Mesh::Point(0.f, 1.f, 0.f));

    m.add_face(vhandle[0], vhandle[1], vhandle[2]);
    m.add_face(vhandle[0], vhandle[2], vhandle[3]);

    typedef deform::AsRigidAsPossibleDeformation< deform::OpenMeshAdapter<> > ARAP;

    deform::OpenMeshAdapter<> adapter(m);

    ARAP arap(adapter);
    REQUIRE(arap.deform(0));

    Eigen::MatrixXf sp = deform::PrivateAccessor<ARAP>::cotanWeights(arap);

    REQUIRE(sp.rows() == 4);
    REQUIRE(sp.cols() == 4);

    Eigen::MatrixXf expected(4, 4);
    expected <<
    0.f, 0.5f, 0.f, 0.5f,
    0.5f, 0.f, 0.5f, 0.f,
    0.f, 0.5f, 0.f, 0.5f,
    0.5f, 0.f, 0.5f, 0.f;

    REQUIRE(sp.isApprox(expected, 1e-4f));
}

// The adapter's five required members and the optional bulk pointer agree with each other, and a deformation driven through
// the adapter (PrecisionType double on a float mesh, as examples/deform_bar.cpp:38 does) moves the OpenMesh points in place.
TEST_CASE("openmesh_adapter_deform")
{
    typedef OpenMesh::TriMesh_ArrayKernelT<> Mesh;
    Mesh m;
    const float v[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const int f[8][3] = {{0, 2, 4}, {2, 1, 4}, {1, 3, 4}, {3, 0, 4}, {2, 0, 5}, {1, 2, 5}, {3, 1, 5}, {0, 3, 5}};
    for (int i = 0; i < 6; ++i) m.add_vertex(Mesh::Point(v[i][0], v[i][1], v[i][2]));
    for (int i = 0; i < 8; ++i) m.add_face(m.vertex_handle(f[i][0]), m.vertex_handle(f[i][1]), m.vertex_handle(f[i][2]));

    typedef deform::OpenMeshAdapter<> Adapter;
    Adapter ma(m);
    REQUIRE(ma.numberOfVertices() == 6);
    REQUIRE(ma.numberOfFaces() == 8);
    for (int i = 0; i < 8; ++i) {
        const Adapter::FaceType ids = ma.face(i);
        REQUIRE(ids(0) == f[i][0]);
        REQUIRE(ids(1) == f[i][1]);
        REQUIRE(ids(2) == f[i][2]);
    }
    const float *bulk = const_cast<const Adapter &>(ma).vertexData();      // packed x,y,z in vertex order
    for (int i = 0; i < 6; ++i)
        for (int d = 0; d < 3; ++d) REQUIRE(bulk[3 * i + d] == ma.vertexLocation(i)(d));

    deform::AsRigidAsPossibleDeformation<Adapter, double> arap(ma);
    arap.setConstraint(5, ma.vertexLocation(5));
    arap.setConstraint(4, Eigen::Vector3f(0.f, 0.f, 1.5f));
    REQUIRE(arap.deform(5));
    REQUIRE(ma.vertexLocation(4).isApprox(Eigen::Vector3f(0.f, 0.f, 1.5f), 1e-6f));
    REQUIRE(m.point(m.vertex_handle(4))[2] == 1.5f);                       // written into the OpenMesh mesh itself
    REQUIRE(ma.vertexLocation(0)(2) > 0.05f);

    // a vertex named twice before deform(): the last location wins (reference arap.h:83 overwrites the map entry)
    arap.setConstraint(4, Eigen::Vector3f(0.f, 0.f, 9.f));
    arap.setConstraint(4, Eigen::Vector3f(0.f, 0.f, 1.25f));
    REQUIRE(arap.deform(1));
    REQUIRE(m.point(m.vertex_handle(4))[2] == 1.25f);

    // a bad index makes THIS deform() fail (the reference would be undefined behaviour) and does not poison the next one
    arap.setConstraint(99, Eigen::Vector3f(0.f, 0.f, 0.f));
    REQUIRE(!arap.deform(1));
    arap.setConstraint(4, Eigen::Vector3f(0.f, 0.f, 1.0f));
    REQUIRE(arap.deform(1));
    REQUIRE(m.point(m.vertex_handle(4))[2] == 1.0f);
}

MINI_TEST_MAIN
