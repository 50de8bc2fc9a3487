// Simplex mesh container + Gmsh MSH 2.2 ASCII reader with the reference's pre-pass
// (reference src/classes/NavierStokes.cpp:7-53) and the geometric boundary-id fallback
// (reference cpp:107-195).  Stands in for dealii::GridIn / Triangulation on one rank.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace nsb_host {

struct Mesh {
  int dim = 0;
  std::vector<double> points;            // [V][dim]
  std::vector<uint32_t> cells;           // [C][dim+1], vertex order as in the file (= deal.II order)
  std::vector<int> cell_tag;             // material id (201)
  std::vector<uint32_t> faces;           // boundary elements [F][dim] (lines / triangles)
  std::vector<int> face_tag;             // boundary id (101..104)
  int64_t n_vertices() const { return dim ? (int64_t)points.size() / dim : 0; }
  int64_t n_cells() const { return dim ? (int64_t)cells.size() / (dim + 1) : 0; }
  int64_t n_faces() const { return dim ? (int64_t)faces.size() / dim : 0; }
};

// throws std::runtime_error("Could not open mesh file: ...") like the reference's AssertThrow (cpp:13-14)
Mesh read_msh(const std::string& path, int dim);
// little-endian binary dump written by tools/msh.py:write_bin (for the multi-million-cell meshes)
Mesh read_bin(const std::string& path);
Mesh read_mesh(const std::string& path, int dim);   // dispatch on the extension

// All boundary faces of the mesh with their boundary id: tagged faces from the file where
// present, boundary id 0 otherwise (deal.II default).  Returns (cell, local face, id).
struct BoundaryFace { int64_t cell; int face; int id; std::array<uint32_t, 3> v; };
std::vector<BoundaryFace> boundary_faces(const Mesh& m);
// reference cpp:133-194: reassign ids geometrically when one of the four expected ids is missing
bool assign_boundary_ids_geometrically(const Mesh& m, std::vector<BoundaryFace>& bf, int inlet, int outlet, int wall,
                                       int cylinder);

// Cell partition for `nparts` ranks (reference cpp:56: GridTools::partition_triangulation = METIS on the face-neighbour graph
// of the cells, PartGraphRecursive for <= 8 parts, Kway beyond, default options).
//   method 0: contiguous chunks of the cell order (default: deterministic, no third-party code on the path)
//   method 1: METIS on the face-dual graph through the header-less libmetis_static.a of the CUDA toolkit (idx_t = int64);
//             layouts then depend on that METIS build, like the reference's depend on its own
std::vector<int32_t> partition_cells(const Mesh& m, int nparts, int method);

// local face -> local vertices, deal.II ReferenceCells::Triangle / Tetrahedron
inline const int* face_vertices(int dim, int f) {
  static const int tri[3][3] = {{0, 1, -1}, {1, 2, -1}, {2, 0, -1}};
  static const int tet[4][3] = {{0, 1, 2}, {1, 0, 3}, {0, 2, 3}, {2, 1, 3}};
  return dim == 2 ? tri[f] : tet[f];
}

}  // namespace nsb_host
