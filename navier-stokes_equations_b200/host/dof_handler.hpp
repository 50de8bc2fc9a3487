// DoF enumeration, support points, boundary DoF sets, Dirichlet constraints and the scalar
// sparsity pattern of FESystem(FE_SimplexP(2)^dim, FE_SimplexP(1)) on one rank, in deal.II's
// numbering.  Replaces
//   dof_handler.distribute_dofs + DoFRenumbering::component_wise   reference NavierStokes.cpp:83-96
//   VectorTools::interpolate_boundary_values -> AffineConstraints    reference cpp:229-253, 617-639
//   DoFTools::make_sparsity_pattern(..., keep_constrained = true)   reference cpp:256-268
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <vector>

#include "mesh.hpp"

namespace nsb_host {

struct DofHandler {
  int dim = 0, dofs_per_cell = 0;
  int64_t n_u = 0, n_p = 0;
  std::vector<uint32_t> cell_dofs;          // [C][dofs_per_cell], FESystem local order
  std::vector<double> support_points;       // [n_dofs][dim]
  std::vector<unsigned char> component;     // [n_dofs] 0..dim-1 velocity, dim pressure
  std::vector<uint32_t> vertex_dof0;        // [V] first DoF (u_0) of each vertex; pressure: vertex_pdof
  std::vector<uint32_t> vertex_pdof;        // [V]
  // edges (sorted vertex pair) -> first DoF
  std::vector<std::pair<uint64_t, uint32_t>> edge_dof0;   // sorted by key = v_lo * V + v_hi
  int64_t n_dofs() const { return n_u + n_p; }
  void distribute(const Mesh& m);
  uint32_t edge_first_dof(uint32_t a, uint32_t b, int64_t V) const;
};

// Pure Dirichlet constraint set x[dof] = value; "first add wins" like repeated
// interpolate_boundary_values calls into one AffineConstraints object.
struct Constraints {
  std::map<uint32_t, double> lines;
  void add_if_new(uint32_t dof, double v) { lines.emplace(dof, v); }
  void to_arrays(std::vector<uint32_t>& d, std::vector<double>& v) const {
    d.clear(); v.clear();
    for (auto& kv : lines) { d.push_back(kv.first); v.push_back(kv.second); }
  }
};

// DoFs on boundary faces with the given id: velocity components (vertices + lines of the face)
// and/or the pressure DoFs of the face vertices.  value(point, component) is evaluated at the
// support point.  Appends to `c` only where no line exists yet.
void interpolate_boundary_values(const Mesh& m, const DofHandler& dh, const std::vector<BoundaryFace>& bf, int id,
                                 const std::function<double(const double*, int)>& value, bool velocity, bool pressure,
                                 Constraints& c);
// std::map overload semantics (later calls overwrite), used by the Newton BC lifting (cpp:1122-1134)
void interpolate_boundary_values_map(const Mesh& m, const DofHandler& dh, const std::vector<BoundaryFace>& bf, int id,
                                     const std::function<double(const double*, int)>& value,
                                     std::map<uint32_t, double>& out);

void make_sparsity_pattern(const DofHandler& dh, std::vector<int64_t>& rowptr, std::vector<uint32_t>& col);

}  // namespace nsb_host
