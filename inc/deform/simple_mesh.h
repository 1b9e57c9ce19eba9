// deform/simple_mesh.h -- a dependency-free indexed triangle mesh, an OBJ reader, and its adapter.
// Not in the reference (which only ships the OpenMesh adapter); it exists so the API can be used and
// tested where OpenMesh is not installed. The adapter implements exactly the mesh concept the solver
// requires (reference inc/deform/openmesh_adapter.h:55,74-113: Scalar, vertexLocation get/set, face,
// numberOfFaces, numberOfVertices) plus the optional bulk accessors the engine can use as a fast path.
#ifndef DEFORM_SIMPLE_MESH_H
#define DEFORM_SIMPLE_MESH_H

#include <deform/detail/linalg.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace deform {

template <class S>
struct SimpleTriMesh {
    typedef S Scalar;
    std::vector<S> positions;   // x,y,z per vertex
    std::vector<int> triangles; // v0,v1,v2 per face

    int addVertex(S x, S y, S z) { positions.push_back(x); positions.push_back(y); positions.push_back(z); return (int)positions.size() / 3 - 1; }
    int addFace(int a, int b, int c) { triangles.push_back(a); triangles.push_back(b); triangles.push_back(c); return (int)triangles.size() / 3 - 1; }

    /** Reads `v` and `f` records of a Wavefront OBJ; vertex index = order of `v` lines, polygons are fanned. */
    bool readObj(const std::string &path) {
        std::ifstream in(path.c_str());
        if (!in) return false;
        positions.clear();
        triangles.clear();
        std::string line;
        while (std::getline(in, line)) {
            if (line.size() > 1 && line[0] == 'v' && line[1] == ' ') {
                std::istringstream ss(line.substr(2));
                double x, y, z;
                if (ss >> x >> y >> z) addVertex((S)x, (S)y, (S)z);
            } else if (line.size() > 1 && line[0] == 'f' && line[1] == ' ') {
                std::istringstream ss(line.substr(2));
                std::vector<int> ids;
                std::string tok;
                while (ss >> tok) ids.push_back(std::atoi(tok.substr(0, tok.find('/')).c_str()) - 1);
                for (size_t k = 1; k + 1 < ids.size(); ++k) addFace(ids[0], ids[k], ids[k + 1]);
            }
        }
        return !positions.empty();
    }
};

template <class S>
class SimpleMeshAdapter {
public:
    typedef SimpleTriMesh<S> Mesh;
    typedef S Scalar;                                   // required
    typedef Eigen::Matrix<Scalar, 3, 1> VertexType;
    typedef Eigen::Matrix<int, 3, 1> FaceType;

    explicit SimpleMeshAdapter(Mesh &mesh) : _mesh(mesh) {}

    VertexType vertexLocation(int idx) const {           // required
        return VertexType(_mesh.positions[3 * (size_t)idx], _mesh.positions[3 * (size_t)idx + 1], _mesh.positions[3 * (size_t)idx + 2]);
    }
    void vertexLocation(int idx, const VertexType &v) {  // required
        _mesh.positions[3 * (size_t)idx] = v(0);
        _mesh.positions[3 * (size_t)idx + 1] = v(1);
        _mesh.positions[3 * (size_t)idx + 2] = v(2);
    }
    FaceType face(int idx) const {                       // required
        FaceType f;
        f(0) = _mesh.triangles[3 * (size_t)idx];
        f(1) = _mesh.triangles[3 * (size_t)idx + 1];
        f(2) = _mesh.triangles[3 * (size_t)idx + 2];
        return f;
    }
    int numberOfFaces() const { return (int)_mesh.triangles.size() / 3; }      // required
    int numberOfVertices() const { return (int)_mesh.positions.size() / 3; }   // required

    // optional bulk accessors (see deform/arap.h): contiguous x,y,z per vertex and v0,v1,v2 per face
    const Scalar *vertexData() const { return _mesh.positions.data(); }
    Scalar *vertexData() { return _mesh.positions.data(); }
    const int *faceData() const { return _mesh.triangles.data(); }

private:
    Mesh &_mesh;
};

}  // namespace deform

#endif
