// TEST DOUBLE, see ../Mesh/TriMesh_ArrayKernelT.hh: OpenMesh::IO::read_mesh for Wavefront OBJ files (`v` and `f`
// records; vertex index = order of the `v` lines, as with the real reader -- SURVEY.md appendix A6).
#ifndef MOCK_OPENMESH_MESHIO_HH
#define MOCK_OPENMESH_MESHIO_HH

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace OpenMesh {
namespace IO {

template <class Mesh>
bool read_mesh(Mesh &mesh, const std::string &path) {
    std::ifstream in(path.c_str());
    if (!in) return false;
    mesh.clear();
    std::string line;
    while (std::getline(in, line)) {
        if (line.size() > 1 && line[0] == 'v' && line[1] == ' ') {
            std::istringstream ss(line.substr(2));
            double x, y, z;
            typedef typename Mesh::Scalar S;
            if (ss >> x >> y >> z) mesh.add_vertex(typename Mesh::Point((S)x, (S)y, (S)z));
        } else if (line.size() > 1 && line[0] == 'f' && line[1] == ' ') {
            std::istringstream ss(line.substr(2));
            std::vector<int> ids;
            std::string tok;
            while (ss >> tok) ids.push_back(std::atoi(tok.substr(0, tok.find('/')).c_str()) - 1);
            for (size_t k = 1; k + 1 < ids.size(); ++k)
                mesh.add_face(mesh.vertex_handle((unsigned)ids[0]), mesh.vertex_handle((unsigned)ids[k]), mesh.vertex_handle((unsigned)ids[k + 1]));
        }
    }
    return mesh.n_vertices() > 0;
}

}  // namespace IO
}  // namespace OpenMesh

#endif
