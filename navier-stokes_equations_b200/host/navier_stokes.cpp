// Implementation of the host-side NavierStokes<dim> mirror.  Control flow follows the reference
// line by line (citations: reference src/classes/NavierStokes.cpp); every numerical kernel of the
// hot path is a call into the CUDA library through the C ABI (include/nsb200.h).
#include "navier_stokes.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <stdexcept>

namespace nsb_host {

namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
NullBuf g_nullbuf;
std::ostream g_null(&g_nullbuf);
}  // namespace

template <int dim>
NavierStokes<dim>::NavierStokes(const std::string& mesh_file_name_, const unsigned int& degree_velocity_,
                                const unsigned int& degree_pressure_, const double deltat_, const double T_,
                                const double Re_, const double U_m_, TimeScheme time_scheme_,
                                NonlinearMethod nonlinear_method_, std::shared_ptr<Function<dim>> inlet_velocity_,
                                std::shared_ptr<Function<dim>> dirichlet_bc_, std::shared_ptr<Function<dim>> forcing_term_,
                                std::shared_ptr<Function<dim>> initial_condition_, bool use_supg_, const RunOptions& opt)
  : options(opt)
  , mpi_size(opt.nranks)
  , mpi_rank(opt.rank)
  , pcout((opt.rank == 0 && opt.verbose) ? std::cout : g_null)
  , pcout_active(opt.rank == 0 && opt.verbose)
  , mesh_file_name(mesh_file_name_)
  , degree_velocity(degree_velocity_)
  , degree_pressure(degree_pressure_)
  , Re(Re_)
  , U_m(U_m_)
  , deltat(deltat_ > 0 ? deltat_ : compute_default_deltat(Re_))
  , T(T_)
  , time_scheme(time_scheme_)
  , nonlinear_method(nonlinear_method_)
  , theta(time_scheme_ == TimeScheme::CrankNicolson ? 0.5 : 1.0)
  , use_supg(use_supg_)
  , inlet_velocity(inlet_velocity_ ? inlet_velocity_ : std::make_shared<InletVelocity<dim>>(H, U_m_, true))
  , dirichlet_bc(dirichlet_bc_ ? dirichlet_bc_ : std::make_shared<ZeroDirichletBC<dim>>())
  , forcing_term(forcing_term_ ? forcing_term_ : std::make_shared<ForcingTerm<dim>>())
  , initial_condition(initial_condition_ ? initial_condition_ : std::make_shared<InitialCondition<dim>>()) {
  if (degree_velocity != 2 || degree_pressure != 1)
    throw std::runtime_error("only the Taylor-Hood pair P2/P1 of the reference's test cases is supported");
}

template <int dim> NavierStokes<dim>::~NavierStokes() {
  if (dev) nsb_destroy(dev);
}

template <int dim> void NavierStokes<dim>::ck(int rc, const char* what) const {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + nsb_last_error(dev));
}

template <int dim>
std::function<double(const double*, int)> NavierStokes<dim>::eval(const std::shared_ptr<Function<dim>>& f) const {
  return [f](const double* x, int comp) {
    Point<dim> p;
    for (int d = 0; d < dim; ++d) p[d] = x[d];
    return f->value(p, (unsigned int)comp);
  };
}

// ------------------------------------------------------------------------------------ setup
template <int dim> void NavierStokes<dim>::setup() {
  pcout << "===============================================" << std::endl;
  pcout << "Setup..." << std::endl;
  mesh = read_mesh(mesh_file_name, dim);                       // cpp:7-53

  double U_mean = (dim == 2) ? (2.0 / 3.0) * U_m : (4.0 / 9.0) * U_m;   // cpp:64-70
  nu = (U_mean * D) / Re;
  pcout << "  Reynolds number: " << Re << std::endl;
  pcout << "  U_max (Inlet param): " << U_m << std::endl;
  pcout << "  U_mean (Reference): " << U_mean << std::endl;
  pcout << "  Cylinder Diameter (D): " << D << std::endl;
  pcout << "  Computed Kinematic viscosity (nu): " << nu << std::endl;
  pcout << "  Time step: " << deltat << std::endl;
  pcout << "  Time scheme: " << to_string(time_scheme) << " (theta=" << theta << ")" << std::endl;
  pcout << "  Nonlinear method: " << to_string(nonlinear_method) << std::endl;

  dof_handler.distribute(mesh);                                // cpp:83-96
  const int64_t n_u = dof_handler.n_u, n_p = dof_handler.n_p;
  pcout << "  Number of active cells: " << mesh.n_cells() << std::endl;
  pcout << "  Number of degrees of freedom: " << dof_handler.n_dofs() << " (" << n_u << " + " << n_p << ")" << std::endl;

  // boundary ids, with the geometric fallback (cpp:107-195)
  bfaces = boundary_faces(mesh);
  {
    std::vector<int> ids;
    for (auto& b : bfaces) ids.push_back(b.id);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    pcout << "  Boundary IDs found in mesh: ";
    for (int id : ids) pcout << id << " ";
    pcout << std::endl;
    pcout << "  Expected IDs: Inlet=" << inlet_boundary_id << ", Walls=" << wall_boundary_id
          << ", Cylinder=" << cylinder_boundary_id << std::endl;
    if (assign_boundary_ids_geometrically(mesh, bfaces, inlet_boundary_id, outlet_boundary_id, wall_boundary_id,
                                          cylinder_boundary_id))
      pcout << "  WARNING: Expected boundary IDs not found! Assigning boundary IDs geometrically..." << std::endl;
  }

  const size_t N = (size_t)dof_handler.n_dofs();                // cpp:205-225
  for (auto* v : {&solution_owned, &solution, &solution_old, &newton_update, &current_solution, &solution_backup,
                  &solution_old_old})
    v->assign(N, 0.0);

  // homogeneous constraints for Newton updates (cpp:229-253)
  {
    auto zero = [](const double*, int) { return 0.0; };
    newton_constraints.lines.clear();
    interpolate_boundary_values(mesh, dof_handler, bfaces, inlet_boundary_id, zero, true, false, newton_constraints);
    interpolate_boundary_values(mesh, dof_handler, bfaces, wall_boundary_id, zero, true, false, newton_constraints);
    interpolate_boundary_values(mesh, dof_handler, bfaces, cylinder_boundary_id, zero, true, false, newton_constraints);
    interpolate_boundary_values(mesh, dof_handler, bfaces, outlet_boundary_id, zero, false, true, newton_constraints);
  }

  // device side: sparsity + matrices (cpp:256-273).  One process per GPU; cells are split into
  // contiguous chunks (the reference: METIS through GridTools::partition_triangulation, cpp:56).
  ck(nsb_create(dim, options.device, &dev), "nsb_create");
  if (!dev) throw std::runtime_error("nsb_create failed (no usable CUDA device)");
  std::vector<int32_t>& part = cell_part;
  part.clear();
  if (mpi_size > 1) {
    ck(nsb_comm_init(dev, (int)mpi_rank, (int)mpi_size, options.nccl_unique_id), "nsb_comm_init");
    part = partition_cells(mesh, (int)mpi_size, options.partitioner);     // cpp:56; output() writes the same ownership
  }
  ck(nsb_set_solver_opts(dev, &options.solver), "nsb_set_solver_opts");
  ck(nsb_upload_mesh(dev, mesh.n_vertices(), mesh.points.data(), mesh.n_cells(), mesh.cells.data(),
                     dof_handler.cell_dofs.data(), n_u, n_p, part.empty() ? nullptr : part.data()),
     "nsb_upload_mesh");
  pressure_matrices_assembled = false;
  pcout << "Setup complete." << std::endl;
}

// The C ABI has no forcing input: every test case of the reference uses the identically-zero ForcingTerm
// (hpp:150-171, TestCases.hpp), and the kernels assemble f = 0 (cpp:683-720).  A user-supplied forcing that is
// not zero would be silently dropped, so it is rejected here instead (sampled at the mesh vertices, current time
// and previous time level).
template <int dim> void NavierStokes<dim>::require_zero_forcing() const {
  if (!forcing_term || forcing_checked_time == forcing_term->get_time()) return;
  forcing_checked_time = forcing_term->get_time();
  const int64_t V = mesh.n_vertices();
  for (int64_t v = 0; v < V; ++v) {
    Point<dim> p;
    for (int d = 0; d < dim; ++d) p[d] = mesh.points[(size_t)v * dim + d];
    for (unsigned int c = 0; c < (unsigned int)dim; ++c)
      if (forcing_term->value(p, c) != 0.0)
        throw std::runtime_error("a non-zero forcing term is not supported by the nsb200 assembly kernels "
                                 "(the reference's test cases all use the zero ForcingTerm)");
  }
}

template <int dim> void NavierStokes<dim>::push_params(bool first_order) {
  require_zero_forcing();
  nsb_params p;
  p.dt = deltat; p.theta = theta; p.nu = nu; p.rho = rho; p.gamma = 0.1;
  p.use_supg = use_supg ? 1 : 0;
  p.first_order_ustar = first_order ? 1 : 0;
  ck(nsb_set_params(dev, &p), "nsb_set_params");
}

// ------------------------------------------------------------------------------------ Newton
template <int dim> void NavierStokes<dim>::assemble_newton_system() {           // cpp:278-539
  std::vector<uint32_t> d;
  std::vector<double> v;
  newton_constraints.to_arrays(d, v);
  ck(nsb_set_constraints(dev, (int64_t)d.size(), d.data(), v.data()), "nsb_set_constraints");
  push_params(true);
  ck(nsb_set_vector(dev, NSB_CURRENT_SOLUTION, current_solution.data()), "nsb_set_vector");
  ck(nsb_set_vector(dev, NSB_SOLUTION_OLD, solution_old.data()), "nsb_set_vector");
  ck(nsb_assemble_newton(dev), "nsb_assemble_newton");
  if (!pressure_matrices_assembled) {
    ck(nsb_assemble_pressure_matrices(dev), "nsb_assemble_pressure_matrices");
    pressure_matrices_assembled = true;
  }
  ck(nsb_rhs_norm(dev, &last_rhs_norm), "nsb_rhs_norm");
}

template <int dim> void NavierStokes<dim>::solve_newton_system() {              // cpp:541-567
  int it = 0;
  double res = 0;
  const int rc = nsb_solve(dev, 500, options.gmres_tolerance, 150, &it, &res);
  ck(rc, "nsb_solve");
  last_gmres_iterations = it;
  step_gmres_iterations += it;
  ++step_solves;
  ck(nsb_get_vector(dev, NSB_SOLUTION, newton_update.data()), "nsb_get_vector");
  const bool forced_fail = options.test_fail_solves > 0 && options.test_fail_solves-- > 0;
  if (rc == 1 || forced_fail) throw NoConvergence(it, res);
  pcout << "  GMRES (Newton): " << it << " iters" << std::endl;
}

// ------------------------------------------------------------------------------------ linearised
template <int dim> void NavierStokes<dim>::build_system_constraints() {         // cpp:617-639
  auto zero = [](const double*, int) { return 0.0; };
  system_constraints.lines.clear();
  interpolate_boundary_values(mesh, dof_handler, bfaces, inlet_boundary_id, eval(inlet_velocity), true, false, system_constraints);
  interpolate_boundary_values(mesh, dof_handler, bfaces, wall_boundary_id, zero, true, false, system_constraints);
  interpolate_boundary_values(mesh, dof_handler, bfaces, cylinder_boundary_id, zero, true, false, system_constraints);
  interpolate_boundary_values(mesh, dof_handler, bfaces, outlet_boundary_id, zero, false, true, system_constraints);
}

template <int dim> void NavierStokes<dim>::assemble_linearized_system() {       // cpp:569-831
  build_system_constraints();
  std::vector<uint32_t> d;
  std::vector<double> v;
  system_constraints.to_arrays(d, v);
  ck(nsb_set_constraints(dev, (int64_t)d.size(), d.data(), v.data()), "nsb_set_constraints");
  push_params(first_step || second_step || time_scheme == TimeScheme::BackwardEuler);   // cpp:665
  ck(nsb_set_vector(dev, NSB_SOLUTION_OLD, solution_old.data()), "nsb_set_vector");
  ck(nsb_set_vector(dev, NSB_SOLUTION_OLD_OLD, solution_old_old.data()), "nsb_set_vector");
  ck(nsb_assemble_linearized(dev), "nsb_assemble_linearized");
  if (!pressure_matrices_assembled) {
    ck(nsb_assemble_pressure_matrices(dev), "nsb_assemble_pressure_matrices");
    pressure_matrices_assembled = true;
  }
}

template <int dim> bool NavierStokes<dim>::solve_linear_system() {              // cpp:833-868
  int it = 0;
  double res = 0;
  const int rc = nsb_solve(dev, 200, options.gmres_tolerance, 150, &it, &res);
  ck(rc, "nsb_solve");
  last_gmres_iterations = it;
  step_gmres_iterations += it;
  ++step_solves;
  const bool forced_fail = options.test_fail_solves > 0 && options.test_fail_solves-- > 0;
  const bool converged = (rc == 0) && !forced_fail;
  if (!converged)
    pcout << "  WARNING: GMRES did NOT converge after " << it << " iterations, residual = " << res << std::endl;
  ck(nsb_get_vector(dev, NSB_SOLUTION, solution_owned.data()), "nsb_get_vector");
  if (converged) pcout << "  GMRES: " << it << " iters" << std::endl;
  return converged;
}

// ------------------------------------------------------------------------------------ post-processing
template <int dim> double NavierStokes<dim>::compute_pressure_difference() {    // cpp:871-912
  double pf[3] = {0.15, 0.2, 0}, pe[3] = {0.25, 0.2, 0};
  if (dim == 3) { pf[0] = 0.205; pf[1] = 0.2; pf[2] = 0.40; pe[0] = 0.205; pe[1] = 0.2; pe[2] = 0.50; }
  const int NV = dim + 1;
  auto evaluate_pressure = [&](const double* pt, bool& found) -> double {
    // VectorTools::point_value: first cell (in cell order) that contains the point
    for (int64_t c = 0; c < mesh.n_cells(); ++c) {
      const uint32_t* cv = &mesh.cells[(size_t)c * NV];
      double X[4][3] = {{0}};
      for (int v = 0; v < NV; ++v)
        for (int k = 0; k < dim; ++k) X[v][k] = mesh.points[(size_t)cv[v] * dim + k];
      // cheap bounding-box reject
      bool out = false;
      for (int k = 0; k < dim && !out; ++k) {
        double lo = X[0][k], hi = X[0][k];
        for (int v = 1; v < NV; ++v) { lo = std::min(lo, X[v][k]); hi = std::max(hi, X[v][k]); }
        out = pt[k] < lo - 1e-10 || pt[k] > hi + 1e-10;
      }
      if (out) continue;
      double lam[4];
      if (dim == 2) {
        const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0], cc = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
        const double det = a * d - b * cc, rx = pt[0] - X[0][0], ry = pt[1] - X[0][1];
        lam[1] = (d * rx - b * ry) / det;
        lam[2] = (-cc * rx + a * ry) / det;
        lam[0] = 1.0 - lam[1] - lam[2];
      } else {
        double J[3][3], r[3];
        for (int i = 0; i < 3; ++i) { r[i] = pt[i] - X[0][i]; for (int k = 0; k < 3; ++k) J[i][k] = X[k + 1][i] - X[0][i]; }
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        auto det3 = [](double a0, double a1, double a2, double b0, double b1, double b2, double c0, double c1, double c2) {
          return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
        };
        lam[1] = det3(r[0], J[0][1], J[0][2], r[1], J[1][1], J[1][2], r[2], J[2][1], J[2][2]) / det;
        lam[2] = det3(J[0][0], r[0], J[0][2], J[1][0], r[1], J[1][2], J[2][0], r[2], J[2][2]) / det;
        lam[3] = det3(J[0][0], J[0][1], r[0], J[1][0], J[1][1], r[1], J[2][0], J[2][1], r[2]) / det;
        lam[0] = 1.0 - lam[1] - lam[2] - lam[3];
      }
      bool inside = true;
      for (int v = 0; v < NV; ++v) inside &= lam[v] >= -1e-10;
      if (!inside) continue;
      double p = 0;
      for (int v = 0; v < NV; ++v) p += lam[v] * current_solution[dof_handler.vertex_pdof[cv[v]]];
      found = true;
      return p;
    }
    found = false;
    return 0.0;
  };
  auto eval_or_warn = [&](const double* pt) {
    bool found = false;
    const double p = evaluate_pressure(pt, found);
    if (!found) pcout << "  WARNING: pressure evaluation point not found by any rank!" << std::endl;
    return found ? p : 0.0;
  };
  return eval_or_warn(pf) - eval_or_warn(pe);
}

template <int dim> void NavierStokes<dim>::compute_lift_drag(double& drag_coeff, double& lift_coeff) const {   // cpp:913-1011
  const int NV = dim + 1, NN = (dim == 2) ? 6 : 10;
  static const int lines2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int lines3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  // QGaussSimplex<dim-1>(degree_velocity + 1): 3-point Gauss on [0,1] / the 7-point triangle rule
  std::vector<std::array<double, 2>> fq;
  std::vector<double> fw;
  if (dim == 2) {
    const double g = std::sqrt(0.6);
    fq = {{0.5 * (1 - g), 0}, {0.5, 0}, {0.5 * (1 + g), 0}};
    fw = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  } else {
    fq = {{0.3333333333330, 0.3333333333330}, {0.7974269853530, 0.1012865073230}, {0.1012865073230, 0.7974269853530},
          {0.1012865073230, 0.1012865073230}, {0.0597158717898, 0.4701420641050}, {0.4701420641050, 0.0597158717898},
          {0.4701420641050, 0.4701420641050}};
    fw = {0.5 * 0.225, 0.5 * 0.125939180545, 0.5 * 0.125939180545, 0.5 * 0.125939180545,
          0.5 * 0.132394152789, 0.5 * 0.132394152789, 0.5 * 0.132394152789};
  }
  double force[3] = {0, 0, 0};
  const int K = dof_handler.dofs_per_cell;
  for (const auto& b : bfaces) {
    if ((unsigned int)b.id != cylinder_boundary_id) continue;
    const uint32_t* cv = &mesh.cells[(size_t)b.cell * NV];
    double X[4][3] = {{0}};
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < dim; ++k) X[v][k] = mesh.points[(size_t)cv[v] * dim + k];
    // grad lambda
    double gl[4][3] = {{0}};
    if (dim == 2) {
      const double a = X[1][0] - X[0][0], bb = X[2][0] - X[0][0], cc = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
      const double det = a * d - bb * cc;
      gl[1][0] = d / det; gl[1][1] = -bb / det; gl[2][0] = -cc / det; gl[2][1] = a / det;
      for (int k = 0; k < 2; ++k) gl[0][k] = -(gl[1][k] + gl[2][k]);
    } else {
      double J[3][3];
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) J[r][k] = X[k + 1][r] - X[0][r];
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                   c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double id = 1.0 / (J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02);
      gl[1][0] = c00 * id; gl[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; gl[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      gl[2][0] = c01 * id; gl[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; gl[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      gl[3][0] = c02 * id; gl[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; gl[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      for (int k = 0; k < 3; ++k) gl[0][k] = -((gl[1][k] + gl[2][k]) + gl[3][k]);
    }
    const int* lv = face_vertices(dim, b.face);
    // outward normal of the cell and the face measure
    double n[3] = {0, 0, 0}, meas;
    if (dim == 2) {
      const double tx = X[lv[1]][0] - X[lv[0]][0], ty = X[lv[1]][1] - X[lv[0]][1];
      meas = std::sqrt(tx * tx + ty * ty);
      n[0] = ty / meas; n[1] = -tx / meas;
    } else {
      double e1[3], e2[3];
      for (int k = 0; k < 3; ++k) { e1[k] = X[lv[1]][k] - X[lv[0]][k]; e2[k] = X[lv[2]][k] - X[lv[0]][k]; }
      n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
      meas = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      for (int k = 0; k < 3; ++k) n[k] /= meas;
    }
    int opp = 0;
    for (int v = 0; v < NV; ++v) {
      bool on = false;
      for (int k = 0; k < dim; ++k) on |= (lv[k] == v);
      if (!on) opp = v;
    }
    double dotp = 0;
    for (int k = 0; k < dim; ++k) dotp += n[k] * (X[opp][k] - X[lv[0]][k]);
    if (dotp > 0) for (int k = 0; k < dim; ++k) n[k] = -n[k];
    const uint32_t* cd = &dof_handler.cell_dofs[(size_t)b.cell * K];
    for (size_t q = 0; q < fw.size(); ++q) {
      // barycentric coordinates of the face quadrature point within the cell
      double lam[4] = {0, 0, 0, 0};
      if (dim == 2) { lam[lv[0]] = 1.0 - fq[q][0]; lam[lv[1]] = fq[q][0]; }
      else { lam[lv[0]] = 1.0 - fq[q][0] - fq[q][1]; lam[lv[1]] = fq[q][0]; lam[lv[2]] = fq[q][1]; }
      double grad_u[3][3] = {{0}}, p = 0;
      for (int a = 0; a < NN; ++a) {
        int i, j;
        if (a < NV) i = j = a;
        else { i = (dim == 2 ? lines2 : lines3)[a - NV][0]; j = (dim == 2 ? lines2 : lines3)[a - NV][1]; }
        double gphi[3];
        for (int k = 0; k < dim; ++k)
          gphi[k] = (a < NV) ? (4.0 * lam[i] - 1.0) * gl[i][k] : 4.0 * (lam[j] * gl[i][k] + lam[i] * gl[j][k]);
        for (int c = 0; c < dim; ++c) {
          const int li = a < NV ? a * (dim + 1) + c : NV * (dim + 1) + (a - NV) * dim + c;
          const double u = current_solution[cd[li]];
          for (int k = 0; k < dim; ++k) grad_u[c][k] += u * gphi[k];
        }
      }
      for (int v = 0; v < NV; ++v) p += lam[v] * current_solution[cd[v * (dim + 1) + dim]];
      const double JxW = fw[q] * meas;
      for (int i = 0; i < dim; ++i) {
        double sn = -p * n[i];
        for (int k = 0; k < dim; ++k) sn += rho * nu * (grad_u[i][k] + grad_u[k][i]) * n[k];
        force[i] += -sn * JxW;
      }
    }
  }
  const double U_mean = (dim == 2) ? (2.0 / 3.0) * U_m : (4.0 / 9.0) * U_m;
  const double ref_area = (dim == 3) ? D * H : D;
  const double den = 0.5 * rho * U_mean * U_mean * ref_area;
  if (dim == 2) { drag_coeff = force[0] / den; lift_coeff = force[1] / den; }
  else { drag_coeff = force[2] / den; lift_coeff = force[1] / den; }
}

// ------------------------------------------------------------------------------------ output
template <int dim> void NavierStokes<dim>::output(const unsigned int time_step) {    // cpp:1013-1042
  if (!options.write_vtu) return;
  // One piece per rank holding that rank's OWNED cells (DataOut::write_vtu_with_pvtu_record, cpp:1037-1041):
  // solution_NNNN.<rank>.vtu with velocity (vector), pressure and subdomain, and solution_NNNN.pvtu on rank 0
  // listing every piece.  Ownership = the cell partition handed to nsb_upload_mesh in setup().
  const int NV = dim + 1;
  const int64_t C = mesh.n_cells();
  std::vector<int64_t> owned;
  for (int64_t c = 0; c < C; ++c)
    if (cell_part.empty() || cell_part[c] == (int32_t)mpi_rank) owned.push_back(c);
  // vertices used by the owned cells, renumbered in order of first use
  std::vector<int64_t> vmap((size_t)mesh.n_vertices(), -1), vlist;
  for (int64_t c : owned)
    for (int k = 0; k < NV; ++k) {
      const uint32_t v = mesh.cells[(size_t)c * NV + k];
      if (vmap[v] < 0) { vmap[v] = (int64_t)vlist.size(); vlist.push_back(v); }
    }
  char name[256];
  std::snprintf(name, sizeof(name), "%ssolution_%04u.%u.vtu", options.output_dir.c_str(), time_step, mpi_rank);
  std::ofstream f(name);
  const int64_t V = (int64_t)vlist.size(), CL = (int64_t)owned.size();
  f << std::setprecision(9);
  f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n";
  f << "<Piece NumberOfPoints=\"" << V << "\" NumberOfCells=\"" << CL << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (int64_t i = 0; i < V; ++i) {
    for (int k = 0; k < 3; ++k) f << (k < dim ? mesh.points[(size_t)vlist[i] * dim + k] : 0.0) << " ";
    f << "\n";
  }
  f << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n";
  for (int64_t c : owned) {
    for (int k = 0; k < NV; ++k) f << vmap[mesh.cells[(size_t)c * NV + k]] << " ";
    f << "\n";
  }
  f << "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n";
  for (int64_t c = 0; c < CL; ++c) f << (c + 1) * NV << "\n";
  f << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
  for (int64_t c = 0; c < CL; ++c) f << (dim == 2 ? 5 : 10) << "\n";
  f << "</DataArray>\n</Cells>\n<PointData Vectors=\"velocity\" Scalars=\"pressure\">\n";
  f << "<DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (int64_t i = 0; i < V; ++i) {
    for (int k = 0; k < 3; ++k) f << (k < dim ? current_solution[dof_handler.vertex_dof0[vlist[i]] + k] : 0.0) << " ";
    f << "\n";
  }
  f << "</DataArray>\n<DataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\">\n";
  for (int64_t i = 0; i < V; ++i) f << current_solution[dof_handler.vertex_pdof[vlist[i]]] << "\n";
  f << "</DataArray>\n</PointData>\n<CellData>\n<DataArray type=\"Float32\" Name=\"subdomain\" format=\"ascii\">\n";
  for (int64_t c = 0; c < CL; ++c) f << mpi_rank << "\n";
  f << "</DataArray>\n</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
  if (mpi_rank == 0) {
    std::snprintf(name, sizeof(name), "%ssolution_%04u.pvtu", options.output_dir.c_str(), time_step);
    std::ofstream p(name);
    p << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n";
    p << "<PPointData Vectors=\"velocity\" Scalars=\"pressure\">\n<PDataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\"/>\n<PDataArray type=\"Float64\" Name=\"pressure\"/>\n</PPointData>\n";
    p << "<PCellData>\n<PDataArray type=\"Float32\" Name=\"subdomain\"/>\n</PCellData>\n<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n";
    for (unsigned int r = 0; r < mpi_size; ++r) {
      char piece[128];
      std::snprintf(piece, sizeof(piece), "solution_%04u.%u.vtu", time_step, r);
      p << "<Piece Source=\"" << piece << "\"/>\n";
    }
    p << "</PUnstructuredGrid>\n</VTKFile>\n";
  }
}

// ------------------------------------------------------------------------------------ run
template <int dim> void NavierStokes<dim>::initialize() {                       // cpp:1045-1071
  setup();
  inlet_velocity->set_time(0.0);
  initial_condition->set_time(0.0);
  auto ic = eval(initial_condition);
  const int64_t N = dof_handler.n_dofs();
  for (int64_t g = 0; g < N; ++g)
    solution_owned[g] = ic(&dof_handler.support_points[(size_t)g * dim], dof_handler.component[g]);
  solution = solution_owned;
  solution_old = solution;
  solution_old_old = solution;
  current_solution = solution;
  first_step = true;
  second_step = true;
  time = 0.0;
  time_step_no = 0;
  if (mpi_rank == 0) {
    forces_file.open(options.output_dir + "forces.txt");
    forces_file << "Time\tCd\tCl\tDeltaP" << std::endl;
  }
  output(time_step_no);
}

template <int dim> StepInfo NavierStokes<dim>::advance() {                      // body of the loop cpp:1073-1323
  StepInfo info;
  time += deltat;
  time_step_no++;
  step_gmres_iterations = 0;
  step_solves = 0;
  const double theta_save = theta;
  if (first_step && time_scheme == TimeScheme::CrankNicolson) {
    theta = 1.0;
    pcout << " (using BE for first step)";
  }
  pcout << "Time step " << time_step_no << " at t=" << time << std::flush;
  inlet_velocity->set_time(time);
  forcing_term->set_time(time);
  {
    Point<dim> p_inlet = (dim == 2) ? Point<dim>(0, H / 2.0) : Point<dim>(0, H / 2.0, H / 2.0);
    const double u_current_real = inlet_velocity->value(p_inlet, (dim == 2 ? 0 : 2));
    const double u_case3_theoretical = U_m * std::sin(M_PI * time / 8.0);
    if (std::abs(u_current_real - u_case3_theoretical) < 1e-4 && time > 0.0) {
      const double u_mean_instant = (dim == 2) ? (2.0 / 3.0 * u_current_real) : (4.0 / 9.0 * u_current_real);
      pcout << "   Instantaneous Re: " << (u_mean_instant * D) / nu << std::endl;
    }
  }
  auto wall_start = std::chrono::high_resolution_clock::now();

  if (nonlinear_method == NonlinearMethod::Newton) {                            // cpp:1116-1207
    {
      auto zero = [](const double*, int) { return 0.0; };
      std::map<uint32_t, double> boundary_values;
      interpolate_boundary_values_map(mesh, dof_handler, bfaces, inlet_boundary_id, eval(inlet_velocity), boundary_values);
      interpolate_boundary_values_map(mesh, dof_handler, bfaces, wall_boundary_id, zero, boundary_values);
      interpolate_boundary_values_map(mesh, dof_handler, bfaces, cylinder_boundary_id, zero, boundary_values);
      solution_owned = current_solution;
      for (const auto& kv : boundary_values) solution_owned[kv.first] = kv.second;
      current_solution = solution_owned;
    }
    double residual_norm = 1e10, previous_residual = 1e10;
    unsigned int newton_iter = 0;
    double damping = 1.0;
    while (residual_norm > newton_tolerance && newton_iter < newton_max_iterations) {
      assemble_newton_system();
      residual_norm = last_rhs_norm;
      pcout << " [" << newton_iter << ": " << residual_norm;
      if (damping < 1.0 - 1e-12) pcout << " a=" << damping;
      pcout << "]" << std::flush;
      if (residual_norm < newton_tolerance) break;
      if (newton_iter > 0 && residual_norm > 0.99 * previous_residual) damping = std::max(0.05, damping * 0.5);
      else if (residual_norm < 0.5 * previous_residual && damping < 1.0 - 1e-12) damping = std::min(1.0, damping * 1.5);
      previous_residual = residual_norm;
      solution_backup = solution_owned;
      bool linear_solve_ok = true;
      try {
        solve_newton_system();
      } catch (const std::exception&) {
        pcout << "(linfail)" << std::flush;
        linear_solve_ok = false;
        damping = std::max(0.05, damping * 0.25);
      }
      solution_owned = current_solution;
      for (size_t i = 0; i < solution_owned.size(); ++i) solution_owned[i] += damping * newton_update[i];
      current_solution = solution_owned;
      if (!linear_solve_ok) {
        assemble_newton_system();
        const double new_res = last_rhs_norm;
        if (new_res > 2.0 * residual_norm) {
          solution_owned = solution_backup;
          current_solution = solution_owned;
          damping = std::max(0.01, damping * 0.5);
          for (size_t i = 0; i < solution_owned.size(); ++i) solution_owned[i] += damping * newton_update[i];
          current_solution = solution_owned;
        }
      }
      newton_iter++;
    }
    info.newton_iterations = (int)newton_iter;
    info.converged = !(residual_norm > newton_tolerance);
    pcout << " Newton: " << newton_iter << " iters, res=" << residual_norm;
    if (residual_norm > newton_tolerance) pcout << " WARNING: Newton did NOT converge!";
    pcout << std::endl;
  } else {                                                                      // cpp:1209-1289
    constexpr int max_substeps = 4;
    std::vector<double> solution_checkpoint = solution_old, solution_old_old_checkpoint = solution_old_old;
    const bool first_step_checkpoint = first_step;
    double dt_attempt = deltat;
    bool step_ok = false;
    int substep = 0;
    while (!step_ok && substep <= max_substeps) {
      if (substep > 0) {
        dt_attempt *= 0.5;
        solution_old = solution_checkpoint;
        solution_old_old = solution_old_old_checkpoint;
        first_step = first_step_checkpoint;
        pcout << "  Retrying with dt=" << dt_attempt << " (attempt " << substep + 1 << ")" << std::endl;
      }
      const double deltat_save = deltat;
      deltat = dt_attempt;
      assemble_linearized_system();
      bool gmres_ok = solve_linear_system();
      if (!gmres_ok && substep == 0) {
        pcout << "  Fallback to BE + 1st-order..." << std::endl;
        const double theta_save_inner = theta;
        const bool fs_save = first_step;
        theta = 1.0;
        first_step = true;
        assemble_linearized_system();
        gmres_ok = solve_linear_system();
        theta = theta_save_inner;
        first_step = fs_save;
      }
      deltat = deltat_save;
      if (gmres_ok) {
        step_ok = true;
        if (substep > 0) pcout << "  Step accepted with reduced dt=" << dt_attempt << std::endl;
      } else {
        substep++;
      }
    }
    if (!step_ok) {
      pcout << "  CRITICAL: all attempts failed. Restoring checkpoint and forcing BE dt=" << dt_attempt << std::endl;
      solution_old = solution_checkpoint;
      solution_old_old = solution_old_old_checkpoint;
      first_step = first_step_checkpoint;
      const double theta_save_inner = theta;
      const bool fs_save = first_step;
      const double deltat_save = deltat;
      theta = 1.0;
      first_step = true;
      deltat = dt_attempt;
      assemble_linearized_system();
      solve_linear_system();
      theta = theta_save_inner;
      first_step = fs_save;
      deltat = deltat_save;
    }
    info.converged = step_ok;
    current_solution = solution_owned;
  }
  {
    auto wall_end = std::chrono::high_resolution_clock::now();
    info.wall_seconds = std::chrono::duration<double>(wall_end - wall_start).count();
    pcout << "  Wall time: " << info.wall_seconds << " s" << std::endl;
  }
  solution_old_old = solution_old;                                               // cpp:1299-1305
  solution_old = current_solution;
  second_step = first_step;
  first_step = false;
  theta = theta_save;

  double drag = 0.0, lift = 0.0;                                                 // cpp:1308-1319
  compute_lift_drag(drag, lift);
  const double delta_p = compute_pressure_difference();
  pcout << "  Cd=" << drag << "  Cl=" << lift << "  dP=" << delta_p << std::endl;
  if (mpi_rank == 0 && forces_file.is_open()) {
    forces_file << time << "\t" << drag << "\t" << lift << "\t" << delta_p << std::endl;
    forces_file.flush();
  }
  output(time_step_no);
  info.time = time; info.cd = drag; info.cl = lift; info.dp = delta_p;
  info.gmres_iterations = step_gmres_iterations;
  info.solves = step_solves;
  return info;
}

template <int dim> void NavierStokes<dim>::run() {                              // cpp:1044-1327
  initialize();
  int steps = 0;
  while (time < T) {
    advance();
    if (options.max_steps >= 0 && ++steps >= options.max_steps) break;
  }
  pcout << "===============================================" << std::endl;
  pcout << "Simulation complete." << std::endl;
}

template class NavierStokes<2>;
template class NavierStokes<3>;

}  // namespace nsb_host
