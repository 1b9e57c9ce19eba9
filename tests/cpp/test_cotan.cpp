// Restatement of reference tests/test_cotan.cpp:20-54 against the B200 facade (needs a GPU).
// Differences forced by the environment: SimpleTriMesh/SimpleMeshAdapter instead of OpenMesh (not installed),
// the bundled Eigen stand-in when Eigen is not installed. The assertions are the reference's.
#include "mini_test.h"

#include <deform/arap.h>
#include <deform/simple_mesh.h>
#include <iostream>

#include "accessor.h"

TEST_CASE("cotan_weights")
{
    typedef deform::SimpleTriMesh<float> Mesh;

    Mesh m;

    m.addVertex(0.f, 0.f, 0.f);
    m.addVertex(1.f, 0.f, 0.f);
    m.addVertex(1.f, 1.f, 0.f);
    m.addVertex(0.f, 1.f, 0.f);

    m.addFace(0, 1, 2);
    m.addFace(0, 2, 3);

    typedef deform::AsRigidAsPossibleDeformation< deform::SimpleMeshAdapter<float> > ARAP;

    deform::SimpleMeshAdapter<float> adapter(m);

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

// The demo call sequence of reference examples/deform_sphere.cpp:55-91 on a small closed mesh (an octahedron):
// pin one vertex, drag the opposite one, deform(5); then the reference's dirty protocol.
TEST_CASE("deform_call_sequence")
{
    typedef deform::SimpleTriMesh<float> Mesh;
    Mesh m;
    m.addVertex(1, 0, 0); m.addVertex(-1, 0, 0); m.addVertex(0, 1, 0); m.addVertex(0, -1, 0); m.addVertex(0, 0, 1); m.addVertex(0, 0, -1);
    const int f[8][3] = {{0, 2, 4}, {2, 1, 4}, {1, 3, 4}, {3, 0, 4}, {2, 0, 5}, {1, 2, 5}, {3, 1, 5}, {0, 3, 5}};
    for (int i = 0; i < 8; ++i) m.addFace(f[i][0], f[i][1], f[i][2]);

    typedef deform::SimpleMeshAdapter<float> Adapter;
    Adapter ma(m);
    deform::AsRigidAsPossibleDeformation<Adapter, double> arap(ma);

    arap.setConstraint(5, ma.vertexLocation(5));
    arap.setConstraint(4, Eigen::Vector3f(0.f, 0.f, 1.5f));
    REQUIRE(arap.deform(5));
    REQUIRE(ma.vertexLocation(4).isApprox(Eigen::Vector3f(0.f, 0.f, 1.5f), 1e-6f));   // handle snapped to its target
    REQUIRE(ma.vertexLocation(5).isApprox(Eigen::Vector3f(0.f, 0.f, -1.f), 1e-6f));   // anchor stays
    REQUIRE(ma.vertexLocation(0)(2) > 0.05f);                                           // the ring follows the handle
    const double e5 = arap.energy();
    REQUIRE(arap.deform(5));                                                            // warm continuation
    REQUIRE(arap.energy() <= e5 * (1 + 1e-9));
}

MINI_TEST_MAIN
