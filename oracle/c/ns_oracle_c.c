/* ORACLE, C port (test infrastructure / CPU baseline -- NOT product code; nothing under
 * navier-stokes_equations_b200/ links or loads this file).
 *
 * PARITY UNPINNED by the reference (no tests / golden vectors; deal.II + Trilinos + MPI absent, so the
 * reference itself cannot be built here).  This file restates, in plain C with OpenMP over cells / rows /
 * row blocks, what the reference executes per time step on the CPU:
 *
 *   nso_assemble_newton       NavierStokes<dim>::assemble_newton_system()      reference src/classes/NavierStokes.cpp:278-539
 *   nso_assemble_linearized   NavierStokes<dim>::assemble_linearized_system()  reference src/classes/NavierStokes.cpp:569-831
 *                             (per-cell, per-q, per-(i,j) loops on vector-valued shape functions exactly as
 *                             written there, tau recomputed inside the loops like cpp:725-729 / 769-773) and
 *                             AffineConstraints::distribute_local_to_global (cpp:810-817)
 *   nso_solve                 solve_linear_system()  cpp:833-868: left-preconditioned GMRES(150), x0 = 0, stop at
 *                             tol_rel*||b||, with PreconditionBlockTriangular (reference NavierStokes.hpp:279-366):
 *                             ILU(1) on F and ILU(0) on M_p as additive Schwarz with zero overlap over `nblocks`
 *                             row blocks (Ifpack's per-MPI-rank ILU; nblocks plays the role of the rank count),
 *                             re-factorized at every solve like hpp:302-307.  K_p^-1 (Trilinos ML in the
 *                             reference, hpp:310-315, not restatable) is a CG solve preconditioned by ILU(0).
 * It is validated against the numpy oracle (tests/test_oracle_c.py) and timed by bench.py as the CPU baseline.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* ------------------------------------------------------------------ finite element tables */
typedef struct {
  int dim, nv, nn, dpc, nq;
  double lam[16][4], w[16];
  int idx[10][2];
  int node[34], comp[34];
} fe_t;

static void fe_init(fe_t *T, int dim) {
  static const int l2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int l3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  memset(T, 0, sizeof(*T));
  T->dim = dim; T->nv = dim + 1;
  const int nl = dim == 2 ? 3 : 6;
  T->nn = T->nv + nl;
  T->dpc = dim * T->nn + T->nv;
  for (int v = 0; v < T->nv; ++v) T->idx[v][0] = T->idx[v][1] = v;
  for (int l = 0; l < nl; ++l) {
    T->idx[T->nv + l][0] = dim == 2 ? l2[l][0] : l3[l][0];
    T->idx[T->nv + l][1] = dim == 2 ? l2[l][1] : l3[l][1];
  }
  int k = 0;
  for (int v = 0; v < T->nv; ++v)
    for (int c = 0; c <= dim; ++c) { T->node[k] = v; T->comp[k] = c; ++k; }
  for (int l = 0; l < nl; ++l)
    for (int c = 0; c < dim; ++c) { T->node[k] = T->nv + l; T->comp[k] = c; ++k; }
  if (dim == 2) {
    static const double p[7][2] = {{0.3333333333330, 0.3333333333330}, {0.7974269853530, 0.1012865073230},
                                   {0.1012865073230, 0.7974269853530}, {0.1012865073230, 0.1012865073230},
                                   {0.0597158717898, 0.4701420641050}, {0.4701420641050, 0.0597158717898},
                                   {0.4701420641050, 0.4701420641050}};
    static const double w[7] = {0.225, 0.125939180545, 0.125939180545, 0.125939180545, 0.132394152789, 0.132394152789, 0.132394152789};
    T->nq = 7;
    for (int q = 0; q < 7; ++q) {
      T->lam[q][0] = 1.0 - (p[q][0] + p[q][1]); T->lam[q][1] = p[q][0]; T->lam[q][2] = p[q][1];
      T->w[q] = 0.5 * w[q];
    }
  } else {
    const double a = 0.5684305841968444, b = 0.1438564719343852;
    const double p[10][3] = {{a, b, b}, {b, b, b}, {b, b, a}, {b, a, b}, {0.0, 0.5, 0.5}, {0.5, 0.0, 0.5}, {0.5, 0.5, 0.0},
                             {0.5, 0.0, 0.0}, {0.0, 0.5, 0.0}, {0.0, 0.0, 0.5}};
    T->nq = 10;
    for (int q = 0; q < 10; ++q) {
      T->lam[q][0] = 1.0 - ((p[q][0] + p[q][1]) + p[q][2]);
      T->lam[q][1] = p[q][0]; T->lam[q][2] = p[q][1]; T->lam[q][3] = p[q][2];
      T->w[q] = (q < 4 ? 0.2177650698804054 : 0.0214899534130631) / 6.0;
    }
  }
}

static int64_t csr_find(const int64_t *rowptr, const int32_t *col, int64_t row, int32_t c) {
  int64_t lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* ------------------------------------------------------------------ assembly (reference loops) */
/* newton == 0: assemble_linearized_system (cpp:569-831), vectors (u^n, u^{n-1});
 * newton == 1: assemble_newton_system (cpp:278-539), vectors (u^k, u^n) -- Jacobian and minus the residual. */
static void assemble_impl(int newton, int dim, int64_t n_cells, const double *points, const int32_t *cells,
                          const int32_t *cell_dofs, int64_t N, const int64_t *rowptr, const int32_t *col,
                          const unsigned char *is_c, const double *cval, const double *sol_old, const double *sol_old_old,
                          double deltat, double theta, double nu, int use_supg, double gamma, int first_order_ustar,
                          double *A, double *b, double *Mp, double *Kp) {
  fe_t T;
  fe_init(&T, dim);
  const int K = T.dpc, NV = T.nv, NQ = T.nq;
  memset(A, 0, sizeof(double) * (size_t)rowptr[N]);
  memset(b, 0, sizeof(double) * (size_t)N);
  if (Mp) memset(Mp, 0, sizeof(double) * (size_t)rowptr[N]);
  if (Kp) memset(Kp, 0, sizeof(double) * (size_t)rowptr[N]);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    double cm[34][34], cmp_[34][34], ckp[34][34], cr[34];
    memset(cm, 0, sizeof(cm)); memset(cmp_, 0, sizeof(cmp_)); memset(ckp, 0, sizeof(ckp)); memset(cr, 0, sizeof(cr));
    const int32_t *cv = cells + cell * NV;
    const int32_t *dofs = cell_dofs + cell * K;
    double X[4][3] = {{0}};
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < dim; ++k) X[v][k] = points[(size_t)cv[v] * dim + k];
    /* fe_values.reinit(cell): affine map */
    double gl[4][3] = {{0}}, det;
    if (dim == 2) {
      const double a = X[1][0] - X[0][0], bb = X[2][0] - X[0][0], cc = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
      det = a * d - bb * cc;
      gl[1][0] = d / det; gl[1][1] = -bb / det; gl[2][0] = -cc / det; gl[2][1] = a / det;
      for (int k = 0; k < 2; ++k) gl[0][k] = -(gl[1][k] + gl[2][k]);
    } else {
      double J[3][3];
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) J[r][k] = X[k + 1][r] - X[0][r];
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                   c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double id = 1.0 / det;
      gl[1][0] = c00 * id; gl[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; gl[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      gl[2][0] = c01 * id; gl[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; gl[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      gl[3][0] = c02 * id; gl[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; gl[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      for (int k = 0; k < 3; ++k) gl[0][k] = -((gl[1][k] + gl[2][k]) + gl[3][k]);
    }
    double h = 0;   /* cell->diameter() */
    for (int v = 0; v < NV; ++v)
      for (int w = v + 1; w < NV; ++w) {
        double s = 0;
        for (int k = 0; k < dim; ++k) s += (X[v][k] - X[w][k]) * (X[v][k] - X[w][k]);
        if (sqrt(s) > h) h = sqrt(s);
      }
    for (int q = 0; q < NQ; ++q) {
      const double JxW = T.w[q] * det;
      double phi_u[34][3], grad_phi_u[34][3][3], div_phi_u[34], phi_p[34], grad_phi_p[34][3];
      memset(phi_u, 0, sizeof(phi_u)); memset(grad_phi_u, 0, sizeof(grad_phi_u));
      memset(div_phi_u, 0, sizeof(div_phi_u)); memset(phi_p, 0, sizeof(phi_p)); memset(grad_phi_p, 0, sizeof(grad_phi_p));
      for (int k = 0; k < K; ++k) {
        const int a = T.node[k], c = T.comp[k];
        if (c < dim) {
          const int i = T.idx[a][0], j = T.idx[a][1];
          double Nv, g[3];
          if (a < NV) {
            Nv = T.lam[q][i] * (2.0 * T.lam[q][i] - 1.0);
            for (int m = 0; m < dim; ++m) g[m] = (4.0 * T.lam[q][i] - 1.0) * gl[i][m];
          } else {
            Nv = 4.0 * T.lam[q][i] * T.lam[q][j];
            for (int m = 0; m < dim; ++m) g[m] = 4.0 * (T.lam[q][j] * gl[i][m] + T.lam[q][i] * gl[j][m]);
          }
          phi_u[k][c] = Nv;
          for (int m = 0; m < dim; ++m) grad_phi_u[k][c][m] = g[m];
          div_phi_u[k] = g[c];
        } else {
          phi_p[k] = T.lam[q][a];
          for (int m = 0; m < dim; ++m) grad_phi_p[k][m] = gl[a][m];
        }
      }
      if (newton) {
        /* current iterate u^k, p^k (sol_old here) and previous time level u^n (sol_old_old here)   cpp:344-375 */
        double u_k[3] = {0}, u_o[3] = {0}, gu_k[3][3] = {{0}}, gu_o[3][3] = {{0}}, p_k = 0, gp_k[3] = {0}, lap_k[3] = {0};
        for (int k = 0; k < K; ++k) {
          const double sk = sol_old[dofs[k]], so = sol_old_old[dofs[k]];
          const int a = T.node[k], c = T.comp[k];
          if (c < dim) {
            u_k[c] += sk * phi_u[k][c];
            u_o[c] += so * phi_u[k][c];
            for (int m = 0; m < dim; ++m) { gu_k[c][m] += sk * grad_phi_u[k][c][m]; gu_o[c][m] += so * grad_phi_u[k][c][m]; }
            /* laplacian of the P2 node function: 4 |g_i|^2 (vertex), 8 g_i.g_j (line) */
            const int i = T.idx[a][0], j = T.idx[a][1];
            double gg = 0;
            for (int m = 0; m < dim; ++m) gg += gl[i][m] * gl[j][m];
            lap_k[c] += sk * (a < NV ? 4.0 : 8.0) * gg;
          } else {
            p_k += sk * phi_p[k];
            for (int m = 0; m < dim; ++m) gp_k[m] += sk * grad_phi_p[k][m];
          }
        }
        double conv_k[3] = {0}, conv_o[3] = {0}, tr = 0, um = 0;
        for (int c = 0; c < dim; ++c) {
          for (int m = 0; m < dim; ++m) { conv_k[c] += gu_k[c][m] * u_k[m]; conv_o[c] += gu_o[c][m] * u_o[m]; }
          tr += gu_k[c][c];
          um += u_k[c] * u_k[c];
        }
        um = sqrt(um);
        const double tau = use_supg ? 1.0 / sqrt(pow(2.0 / deltat, 2) + pow(2.0 * um / h, 2) + pow(4.0 * nu / (h * h), 2)) : 0.0;
        double Gu[34][3], gukphi[34][3];
        for (int k = 0; k < K; ++k)
          for (int c = 0; c < dim; ++c) {
            double s1 = 0, s2 = 0;
            for (int m = 0; m < dim; ++m) { s1 += grad_phi_u[k][c][m] * u_k[m]; s2 += gu_k[c][m] * phi_u[k][m]; }
            Gu[k][c] = s1; gukphi[k][c] = s2;
          }
        for (int i = 0; i < K; ++i) {
          /* minus the residual (cpp:377-418) */
          double time_term = 0, conv_impl = 0, visc_impl = 0, conv_expl = 0, visc_expl = 0;
          for (int c = 0; c < dim; ++c) {
            time_term += (u_k[c] - u_o[c]) * phi_u[i][c] / deltat;
            conv_impl += theta * conv_k[c] * phi_u[i][c];
            conv_expl += (1.0 - theta) * conv_o[c] * phi_u[i][c];
            for (int m = 0; m < dim; ++m) {
              visc_impl += theta * nu * gu_k[c][m] * grad_phi_u[i][c][m];
              visc_expl += (1.0 - theta) * nu * gu_o[c][m] * grad_phi_u[i][c][m];
            }
          }
          const double pres_term = -p_k * div_phi_u[i], div_term = -phi_p[i] * tr;
          cr[i] += (-time_term - conv_impl - visc_impl - conv_expl - visc_expl - pres_term - div_term) * JxW;
          if (use_supg) {
            double s = 0;                                   /* cpp:488-509 */
            for (int c = 0; c < dim; ++c)
              s += Gu[i][c] * ((u_k[c] - u_o[c]) / deltat + conv_k[c] + gp_k[c] - nu * lap_k[c]);
            cr[i] -= s * tau * JxW;
          }
          /* Jacobian (cpp:421-466) */
          for (int j = 0; j < K; ++j) {
            double val = 0, visc = 0, cv = 0;
            for (int c = 0; c < dim; ++c) {
              val += phi_u[i][c] * phi_u[j][c] / deltat;
              for (int m = 0; m < dim; ++m) visc += grad_phi_u[i][c][m] * grad_phi_u[j][c][m];
              cv += (Gu[j][c] + gukphi[j][c]) * phi_u[i][c];
            }
            val += theta * nu * visc + theta * cv;
            val -= phi_p[j] * div_phi_u[i];
            val -= phi_p[i] * div_phi_u[j];
            cm[i][j] += val * JxW;
            if (use_supg) {
              double s = 0;
              for (int c = 0; c < dim; ++c)
                s += tau * Gu[i][c] * (phi_u[j][c] / deltat + Gu[j][c] + gukphi[j][c] + grad_phi_p[j][c]);
              cm[i][j] += s * JxW;
              cm[i][j] += gamma * (div_phi_u[i] * div_phi_u[j]) * JxW;
            }
            if (Mp) cmp_[i][j] += phi_p[i] * phi_p[j] * JxW;
            if (Kp) {
              double g = 0;
              for (int m = 0; m < dim; ++m) g += grad_phi_p[i][m] * grad_phi_p[j][m];
              ckp[i][j] += g * JxW;
            }
          }
        }
        continue;
      }
      /* get_function_values / gradients (cpp:653-658) */
      double u_old[3] = {0}, u_oo[3] = {0}, gu_old[3][3] = {{0}};
      for (int k = 0; k < K; ++k) {
        const double so = sol_old[dofs[k]], soo = sol_old_old[dofs[k]];
        for (int c = 0; c < dim; ++c) {
          u_old[c] += so * phi_u[k][c];
          u_oo[c] += soo * phi_u[k][c];
          for (int m = 0; m < dim; ++m) gu_old[c][m] += so * grad_phi_u[k][c][m];
        }
      }
      double u_star[3];
      if (first_order_ustar) {
        for (int c = 0; c < dim; ++c) u_star[c] = u_old[c];
      } else {
        double ns = 0, no = 0;
        for (int c = 0; c < dim; ++c) { u_star[c] = 2.0 * u_old[c] - u_oo[c]; ns += u_star[c] * u_star[c]; no += u_old[c] * u_old[c]; }
        ns = sqrt(ns); no = sqrt(no);
        if (no > 1e-12 && ns > 1.2 * no)
          for (int c = 0; c < dim; ++c) u_star[c] = u_old[c];
      }
      for (int i = 0; i < K; ++i) {
        double rhs_mass = 0, rhs_visc = 0, rhs_conv = 0;
        for (int c = 0; c < dim; ++c) {
          rhs_mass += (1.0 / deltat) * u_old[c] * phi_u[i][c];
          double conv = 0;
          for (int m = 0; m < dim; ++m) { rhs_visc += gu_old[c][m] * grad_phi_u[i][c][m]; conv += gu_old[c][m] * u_old[m]; }
          rhs_conv += conv * phi_u[i][c];
        }
        cr[i] += (rhs_mass - (1.0 - theta) * nu * rhs_visc - (1.0 - theta) * rhs_conv) * JxW;
        if (use_supg) {
          double um = 0;
          for (int c = 0; c < dim; ++c) um += u_star[c] * u_star[c];
          um = sqrt(um);
          const double tau = 1.0 / sqrt(pow(2.0 / deltat, 2) + pow(2.0 * um / h, 2) + pow(4.0 * nu / (h * h), 2));
          double s = 0;   /* (tau * (u_star * grad_phi_u[i])) . (u_old/dt): first-index contraction, cpp:733 */
          for (int m = 0; m < dim; ++m) {
            double t = 0;
            for (int c = 0; c < dim; ++c) t += u_star[c] * grad_phi_u[i][c][m];
            s += tau * t * (u_old[m] / deltat);
          }
          cr[i] += s * JxW;
        }
        for (int j = 0; j < K; ++j) {
          double val = 0, visc = 0, conv = 0;
          for (int c = 0; c < dim; ++c) {
            val += (1.0 / deltat) * phi_u[i][c] * phi_u[j][c];
            double gj_u = 0;
            for (int m = 0; m < dim; ++m) { visc += grad_phi_u[i][c][m] * grad_phi_u[j][c][m]; gj_u += grad_phi_u[j][c][m] * u_star[m]; }
            conv += gj_u * phi_u[i][c];
          }
          val += theta * nu * visc + theta * conv;
          val -= phi_p[j] * div_phi_u[i];
          val -= phi_p[i] * div_phi_u[j];
          cm[i][j] += val * JxW;
          if (use_supg) {
            double um = 0;
            for (int c = 0; c < dim; ++c) um += u_star[c] * u_star[c];
            um = sqrt(um);
            const double tau = 1.0 / sqrt(pow(2.0 / deltat, 2) + pow(2.0 * um / h, 2) + pow(4.0 * nu / (h * h), 2));
            double s1 = 0, s2 = 0;
            for (int c = 0; c < dim; ++c) {
              double gi_u = 0, gj_u = 0;
              for (int m = 0; m < dim; ++m) { gi_u += grad_phi_u[i][c][m] * u_star[m]; gj_u += grad_phi_u[j][c][m] * u_star[m]; }
              s1 += tau * gi_u * (phi_u[j][c] / deltat + gj_u);
              s2 += tau * gi_u * grad_phi_p[j][c];
            }
            cm[i][j] += s1 * JxW;
            cm[i][j] += s2 * JxW;
            cm[i][j] += gamma * (div_phi_u[i] * div_phi_u[j]) * JxW;
          }
          if (Mp) cmp_[i][j] += phi_p[i] * phi_p[j] * JxW;
          if (Kp) {
            double g = 0;
            for (int m = 0; m < dim; ++m) g += grad_phi_p[i][m] * grad_phi_p[j][m];
            ckp[i][j] += g * JxW;
          }
        }
      }
    }
    /* distribute_local_to_global with Dirichlet lines (A.5) */
    double avg = 0, avgm = 0, avgk = 0;
    for (int k = 0; k < K; ++k) { avg += fabs(cm[k][k]); avgm += fabs(cmp_[k][k]); avgk += fabs(ckp[k][k]); }
    avg /= K; avgm /= K; avgk /= K;
    for (int i = 0; i < K; ++i) {
      const int64_t I = dofs[i];
      if (is_c[I]) {
        const int64_t p = csr_find(rowptr, col, I, (int32_t)I);
        const double d = cm[i][i] != 0.0 ? fabs(cm[i][i]) : avg;
#pragma omp atomic
        A[p] += d;
        if (Mp) {
          const double dm = cmp_[i][i] != 0.0 ? fabs(cmp_[i][i]) : avgm;
#pragma omp atomic
          Mp[p] += dm;
        }
        if (Kp) {
          const double dk = ckp[i][i] != 0.0 ? fabs(ckp[i][i]) : avgk;
#pragma omp atomic
          Kp[p] += dk;
        }
        continue;
      }
      double r = cr[i];
      for (int j = 0; j < K; ++j) {
        const int64_t Jd = dofs[j];
        if (is_c[Jd]) { r -= cm[i][j] * cval[Jd]; continue; }
        const int64_t p = csr_find(rowptr, col, I, (int32_t)Jd);
#pragma omp atomic
        A[p] += cm[i][j];
        if (Mp) {
#pragma omp atomic
          Mp[p] += cmp_[i][j];
        }
        if (Kp) {
#pragma omp atomic
          Kp[p] += ckp[i][j];
        }
      }
#pragma omp atomic
      b[I] += r;
    }
  }
  if (Kp && Mp)
    for (int64_t k = 0; k < rowptr[N]; ++k) Kp[k] += 1e-6 * Mp[k];     /* cpp:536, 828 */
}

void nso_assemble_linearized(int dim, int64_t n_cells, const double *points, const int32_t *cells, const int32_t *cell_dofs,
                             int64_t N, const int64_t *rowptr, const int32_t *col, const unsigned char *is_c,
                             const double *cval, const double *sol_old, const double *sol_old_old, double deltat,
                             double theta, double nu, int use_supg, double gamma, int first_order_ustar, double *A,
                             double *b, double *Mp, double *Kp) {
  assemble_impl(0, dim, n_cells, points, cells, cell_dofs, N, rowptr, col, is_c, cval, sol_old, sol_old_old, deltat, theta, nu,
                use_supg, gamma, first_order_ustar, A, b, Mp, Kp);
}

void nso_assemble_newton(int dim, int64_t n_cells, const double *points, const int32_t *cells, const int32_t *cell_dofs,
                         int64_t N, const int64_t *rowptr, const int32_t *col, const unsigned char *is_c, const double *cval,
                         const double *sol_current, const double *sol_old, double deltat, double theta, double nu, int use_supg,
                         double gamma, double *A, double *b, double *Mp, double *Kp) {
  assemble_impl(1, dim, n_cells, points, cells, cell_dofs, N, rowptr, col, is_c, cval, sol_current, sol_old, deltat, theta, nu,
                use_supg, gamma, 1, A, b, Mp, Kp);
}

/* ------------------------------------------------------------------ ILU(k) over row blocks */
typedef struct {
  int n, nblocks;
  int *bstart;          /* [nblocks+1] */
  int64_t *ptr;         /* [n+1] local-column factor rows, L (unit) and U interleaved, sorted */
  int *col;             /* block-local column */
  double *val;
  int64_t *diag;        /* position of the diagonal in each row */
} ilu_t;

static void ilu_free(ilu_t *F) {
  free(F->bstart); free(F->ptr); free(F->col); free(F->val); free(F->diag);
  memset(F, 0, sizeof(*F));
}

/* A: n x n CSR (int64 ptr, int32 col).  Factor each diagonal block [bstart[k], bstart[k+1]) with level-of-fill `lof`. */
static void ilu_setup(ilu_t *F, int n, const int64_t *ptr, const int32_t *col, const double *val, int lof, int nblocks) {
  memset(F, 0, sizeof(*F));
  F->n = n; F->nblocks = nblocks;
  F->bstart = (int *)malloc(sizeof(int) * (nblocks + 1));
  for (int k = 0; k <= nblocks; ++k) F->bstart[k] = (int)((int64_t)n * k / nblocks);
  int64_t **rptr = (int64_t **)calloc(nblocks, sizeof(int64_t *));
  int **rcol = (int **)calloc(nblocks, sizeof(int *));
  double **rval = (double **)calloc(nblocks, sizeof(double *));
  int64_t **rdiag = (int64_t **)calloc(nblocks, sizeof(int64_t *));
#pragma omp parallel for schedule(dynamic, 1)
  for (int blk = 0; blk < nblocks; ++blk) {
    const int r0 = F->bstart[blk], r1 = F->bstart[blk + 1], nb = r1 - r0;
    int64_t cap = 0;
    for (int i = r0; i < r1; ++i) cap += ptr[i + 1] - ptr[i];
    cap = cap * (lof > 0 ? 3 : 1) + 16;
    int64_t *bp = (int64_t *)malloc(sizeof(int64_t) * (nb + 1));
    int *bc = (int *)malloc(sizeof(int) * cap);
    int *bl = (int *)malloc(sizeof(int) * cap);        /* levels */
    double *bv = (double *)malloc(sizeof(double) * cap);
    int64_t *bd = (int64_t *)malloc(sizeof(int64_t) * nb);
    int *lev = (int *)malloc(sizeof(int) * nb);
    double *w = (double *)calloc(nb, sizeof(double));
    int *list = (int *)malloc(sizeof(int) * nb);
    for (int i = 0; i < nb; ++i) lev[i] = -1;
    bp[0] = 0;
    for (int li = 0; li < nb; ++li) {
      const int gi = r0 + li;
      int cnt = 0;
      for (int64_t k = ptr[gi]; k < ptr[gi + 1]; ++k) {
        const int c = col[k] - r0;
        if (c < 0 || c >= nb) continue;                /* zero overlap: couplings to other blocks are dropped */
        lev[c] = 0; w[c] = val[k]; list[cnt++] = c;
      }
      if (lev[li] < 0) { lev[li] = 0; w[li] = 0.0; list[cnt++] = li; }
      /* keep the list sorted; eliminate with previous rows in increasing column order (fill has larger columns) */
      for (int a = 1; a < cnt; ++a) {
        const int x = list[a];
        int bpos = a - 1;
        while (bpos >= 0 && list[bpos] > x) { list[bpos + 1] = list[bpos]; --bpos; }
        list[bpos + 1] = x;
      }
      for (int t = 0; t < cnt && list[t] < li; ++t) {
        const int kcol = list[t];
        const int klev = lev[kcol];
        const double mult = w[kcol] / bv[bd[kcol]];
        w[kcol] = mult;
        int ins = t + 1;                                  /* U row kcol is sorted too: merge position only moves forward */
        for (int64_t p = bd[kcol] + 1; p < bp[kcol + 1]; ++p) {
          const int c = bc[p];
          const int nl = klev + bl[p] + 1;
          if (lev[c] == -1) {
            if (nl > lof) continue;
            while (ins < cnt && list[ins] < c) ++ins;
            memmove(list + ins + 1, list + ins, sizeof(int) * (cnt - ins));
            list[ins] = c; ++cnt;
            lev[c] = nl; w[c] = 0.0;
          } else if (nl < lev[c]) lev[c] = nl;
          w[c] -= mult * bv[p];
        }
      }
      if (bp[li] + cnt > cap) {
        cap = (bp[li] + cnt) * 2;
        bc = (int *)realloc(bc, sizeof(int) * cap); bl = (int *)realloc(bl, sizeof(int) * cap); bv = (double *)realloc(bv, sizeof(double) * cap);
      }
      int64_t p = bp[li];
      for (int t = 0; t < cnt; ++t) {
        const int c = list[t];
        bc[p] = c; bv[p] = w[c];
        bl[p] = lev[c];
        if (c == li) { bd[li] = p; if (bv[p] == 0.0) bv[p] = 1e-300; }
        lev[c] = -1; w[c] = 0.0;
        ++p;
      }
      bp[li + 1] = p;
    }
    free(lev); free(w); free(list); free(bl);
    rptr[blk] = bp; rcol[blk] = bc; rval[blk] = bv; rdiag[blk] = bd;
  }
  /* concatenate */
  F->ptr = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  F->diag = (int64_t *)malloc(sizeof(int64_t) * n);
  int64_t tot = 0;
  for (int blk = 0; blk < nblocks; ++blk) tot += rptr[blk][F->bstart[blk + 1] - F->bstart[blk]];
  F->col = (int *)malloc(sizeof(int) * tot);
  F->val = (double *)malloc(sizeof(double) * tot);
  int64_t off = 0;
  for (int blk = 0; blk < nblocks; ++blk) {
    const int r0 = F->bstart[blk], nb = F->bstart[blk + 1] - r0;
    for (int i = 0; i < nb; ++i) { F->ptr[r0 + i] = off + rptr[blk][i]; F->diag[r0 + i] = off + rdiag[blk][i]; }
    memcpy(F->col + off, rcol[blk], sizeof(int) * rptr[blk][nb]);
    memcpy(F->val + off, rval[blk], sizeof(double) * rptr[blk][nb]);
    off += rptr[blk][nb];
    free(rptr[blk]); free(rcol[blk]); free(rval[blk]); free(rdiag[blk]);
  }
  F->ptr[n] = off;
  free(rptr); free(rcol); free(rval); free(rdiag);
}

static void ilu_apply(const ilu_t *F, const double *x, double *y) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int blk = 0; blk < F->nblocks; ++blk) {
    const int r0 = F->bstart[blk], r1 = F->bstart[blk + 1];
    for (int i = r0; i < r1; ++i) {                      /* L z = x (unit lower) */
      double s = x[i];
      for (int64_t p = F->ptr[i]; p < F->diag[i]; ++p) s -= F->val[p] * y[r0 + F->col[p]];
      y[i] = s;
    }
    for (int i = r1 - 1; i >= r0; --i) {                 /* U y = z */
      double s = y[i];
      for (int64_t p = F->diag[i] + 1; p < F->ptr[i + 1]; ++p) s -= F->val[p] * y[r0 + F->col[p]];
      y[i] = s / F->val[F->diag[i]];
    }
  }
}

/* ------------------------------------------------------------------ small CSR helpers */
typedef struct { int n; int64_t *ptr; int32_t *col; double *val; } csr_t;

static void csr_block(const int64_t *rowptr, const int32_t *col, const double *val, int64_t r0, int64_t r1, int64_t c0, int64_t c1, csr_t *B) {
  const int n = (int)(r1 - r0);
  B->n = n;
  B->ptr = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  int64_t cnt = 0;
  for (int64_t i = r0; i < r1; ++i)
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) cnt += (col[k] >= c0 && col[k] < c1);
  B->col = (int32_t *)malloc(sizeof(int32_t) * (cnt + 1));
  B->val = (double *)malloc(sizeof(double) * (cnt + 1));
  cnt = 0;
  for (int64_t i = r0; i < r1; ++i) {
    B->ptr[i - r0] = cnt;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] >= c0 && col[k] < c1) { B->col[cnt] = (int32_t)(col[k] - c0); B->val[cnt] = val[k]; ++cnt; }
  }
  B->ptr[n] = cnt;
}
static void csr_free(csr_t *B) { free(B->ptr); free(B->col); free(B->val); }
static void csr_mv(const csr_t *B, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < B->n; ++i) {
    double s = 0;
    for (int64_t k = B->ptr[i]; k < B->ptr[i + 1]; ++k) s += B->val[k] * x[B->col[k]];
    y[i] = s;
  }
}
static double dotp(int64_t n, const double *a, const double *b) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* K_p^-1 t by ILU(0)-preconditioned CG.  rtol must be far below the outer tolerance: CG is not a fixed linear
 * operator, and plain (non-flexible) GMRES needs one. */
static void kp_solve(const csr_t *Kp, const ilu_t *Fk, const double *t, double *x, double *wk, double rtol) {
  const int n = Kp->n;
  double *r = wk, *z = wk + n, *p = wk + 2 * n, *q = wk + 3 * n;
  memset(x, 0, sizeof(double) * n);
  memcpy(r, t, sizeof(double) * n);
  const double r0 = sqrt(dotp(n, r, r));
  if (r0 == 0 || r0 != r0) return;
  ilu_apply(Fk, r, z);
  memcpy(p, z, sizeof(double) * n);
  double rz = dotp(n, r, z);
  for (int it = 0; it < 1000; ++it) {
    csr_mv(Kp, p, q);
    const double alpha = rz / dotp(n, p, q);
    for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
    const double rn_ = sqrt(dotp(n, r, r));
    if (rn_ <= rtol * r0 || rn_ != rn_) break;
    ilu_apply(Fk, r, z);
    const double rz2 = dotp(n, r, z);
    const double beta = rz2 / rz;
    rz = rz2;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
}

/* ------------------------------------------------------------------ solve_linear_system */
/* Core: A as full scalar CSR (used in place, not copied), the (1,1) blocks of M_p / K_p as a compact CSR over the pressure
 * DoFs (shared pattern pp_ptr / pp_col).  time_budget_s > 0: stop iterating once the solve has run that long (the result is
 * then reported as not converged; bench.py's reference arm uses it to stay inside its time limit).
 * Returns 0 converged, 1 max_it reached, 2 stopped by the time budget, 3 the residual became NaN (ILU breakdown).
 * timings[4] (optional) = {preconditioner setup, GMRES iterations, K_p CG inside them, total} in seconds. */
int nso_solve_blocks(int64_t N, int64_t n_u, const int64_t *rowptr, const int32_t *col, const double *A, const int64_t *pp_ptr,
                     const int32_t *pp_col, const double *Mp, const double *Kp, const double *b, double nu, double rho,
                     double deltat, double theta, int max_it, double tol_rel, int n_tmp_vectors, int nblocks,
                     double schur_mass_coeff, double kp_tol, double time_budget_s, double *x, int *iterations, double *residual,
                     double *timings) {
  /* schur_mass_coeff < 0: the reference's theta*nu (hpp:342); >= 0: override (the product's default is theta*nu + gamma) */
  const double cmass = schur_mass_coeff >= 0 ? schur_mass_coeff : theta * nu;
  const int64_t n_p = N - n_u;
  if (nblocks < 1) nblocks = 1;
  csr_t Af, F, B, Mpp, Kpp;
  Af.n = (int)N; Af.ptr = (int64_t *)rowptr; Af.col = (int32_t *)col; Af.val = (double *)A;     /* borrowed */
  Mpp.n = Kpp.n = (int)n_p;
  Mpp.ptr = Kpp.ptr = (int64_t *)pp_ptr; Mpp.col = Kpp.col = (int32_t *)pp_col;
  Mpp.val = (double *)Mp; Kpp.val = (double *)Kp;
  const double t_setup0 = omp_get_wtime();
  /* PreconditionBlockTriangular::initialize -- every solve (hpp:282-318) */
  ilu_t iF, iM, iK;
  csr_block(rowptr, col, A, 0, n_u, 0, n_u, &F);
  ilu_setup(&iF, F.n, F.ptr, F.col, F.val, 1, nblocks);
  const long long nnzF = (long long)F.ptr[F.n];
  csr_free(&F);                                          /* only the factors are needed from here on */
  csr_block(rowptr, col, A, n_u, N, 0, n_u, &B);
  ilu_setup(&iM, Mpp.n, Mpp.ptr, Mpp.col, Mpp.val, 0, nblocks);
  ilu_setup(&iK, Kpp.n, Kpp.ptr, Kpp.col, Kpp.val, 0, nblocks);
  const double t_setup = omp_get_wtime() - t_setup0;
  if (getenv("NSO_DEBUG")) fprintf(stderr, "[nso] ilu setup %.3f s, nnz(F)=%lld nnz(ILU1)=%lld\n", t_setup, nnzF, (long long)iF.ptr[iF.n]);
  const int m = n_tmp_vectors - 2 > 1 ? n_tmp_vectors - 2 : 1;
  double *V = (double *)malloc(sizeof(double) * (size_t)(m + 1) * N);      /* pages are touched as the basis grows */
  double *w = (double *)malloc(sizeof(double) * N), *tmp = (double *)malloc(sizeof(double) * N);
  double *tp = (double *)malloc(sizeof(double) * n_p), *t2 = (double *)malloc(sizeof(double) * n_p), *y1 = (double *)malloc(sizeof(double) * n_p);
  double *wk = (double *)malloc(sizeof(double) * 4 * n_p);
  double *H = (double *)calloc((size_t)(m + 1) * m, sizeof(double)), *cs = (double *)calloc(m, sizeof(double)),
         *sn = (double *)calloc(m, sizeof(double)), *g = (double *)calloc(m + 1, sizeof(double)), *yv = (double *)calloc(m, sizeof(double));
#define PRECOND(in, out)                                                                                     \
  do {                                                                                                       \
    ilu_apply(&iF, (in), (out));                        /* y0 = ILU_F^-1 x0        (hpp:325) */               \
    csr_mv(&B, (out), tp);                              /* tmp = B y0              (hpp:334) */               \
    for (int64_t i_ = 0; i_ < n_p; ++i_) tp[i_] = (in)[n_u + i_] - tp[i_];                                    \
    const double tk_ = omp_get_wtime();                                                                      \
    kp_solve(&Kpp, &iK, tp, y1, wk, kp_tol);            /* K_p^-1                  (hpp:338) */               \
    t_kp += omp_get_wtime() - tk_;                                                                           \
    ilu_apply(&iM, tp, t2);                             /* ILU_Mp^-1               (hpp:342) */               \
    for (int64_t i_ = 0; i_ < n_p; ++i_) (out)[n_u + i_] = -(rho / deltat) * y1[i_] - cmass * t2[i_];         \
  } while (0)
  double t_kp = 0;
  memset(x, 0, sizeof(double) * N);
  const double tol = tol_rel * sqrt(dotp(N, b, b));
  int it = 0, rc = 1, first = 1, out_of_time = 0;
  double res = 0;
  for (;;) {
    if (first) { PRECOND(b, w); }
    else {
      csr_mv(&Af, x, tmp);
      for (int64_t i = 0; i < N; ++i) tmp[i] = b[i] - tmp[i];
      PRECOND(tmp, w);
    }
    const double beta = sqrt(dotp(N, w, w));
    res = beta;
    if (first && beta <= tol) { rc = 0; break; }
    first = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) V[i] = w[i] / beta;
    memset(g, 0, sizeof(double) * (m + 1));
    g[0] = beta;
    int kused = 0, conv = 0;
    for (int k = 0; k < m; ++k) {
      double *vk = V + (size_t)k * N, *vn = V + (size_t)(k + 1) * N;
      csr_mv(&Af, vk, tmp);
      PRECOND(tmp, vn);
      for (int i = 0; i <= k; ++i) {                     /* modified Gram-Schmidt */
        const double hik = dotp(N, vn, V + (size_t)i * N);
        H[(size_t)i * m + k] = hik;
        const double *vi = V + (size_t)i * N;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < N; ++j) vn[j] -= hik * vi[j];
      }
      const double hn = sqrt(dotp(N, vn, vn));
      H[(size_t)(k + 1) * m + k] = hn;
      if (hn > 0) {
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < N; ++j) vn[j] /= hn;
      }
      for (int i = 0; i < k; ++i) {
        const double t = cs[i] * H[(size_t)i * m + k] + sn[i] * H[(size_t)(i + 1) * m + k];
        H[(size_t)(i + 1) * m + k] = -sn[i] * H[(size_t)i * m + k] + cs[i] * H[(size_t)(i + 1) * m + k];
        H[(size_t)i * m + k] = t;
      }
      const double a = H[(size_t)k * m + k], bb = H[(size_t)(k + 1) * m + k], d = hypot(a, bb);
      cs[k] = a / d; sn[k] = bb / d;
      H[(size_t)k * m + k] = d; H[(size_t)(k + 1) * m + k] = 0;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      res = fabs(g[k + 1]);
      ++it; kused = k + 1;
      if (getenv("NSO_DEBUG") && it % 10 == 0) fprintf(stderr, "[nso] it %d res %.4e (tol %.4e)\n", it, res, tol);
      if (res <= tol) { conv = 1; break; }
      if (it >= max_it) break;
      if (res != res) { out_of_time = 2; break; }         /* NaN: SolverControl::check reports failure at once (deal.II) */
      if (time_budget_s > 0 && omp_get_wtime() - t_setup0 > time_budget_s) { out_of_time = 1; break; }
    }
    for (int i = kused - 1; i >= 0; --i) {
      double s = g[i];
      for (int j = i + 1; j < kused; ++j) s -= H[(size_t)i * m + j] * yv[j];
      yv[i] = s / H[(size_t)i * m + i];
    }
    for (int i = 0; i < kused; ++i) {
      const double *vi = V + (size_t)i * N;
#pragma omp parallel for schedule(static)
      for (int64_t j = 0; j < N; ++j) x[j] += yv[i] * vi[j];
    }
    if (conv) { rc = 0; break; }
    if (it >= max_it || out_of_time) break;
  }
  const double t_total = omp_get_wtime() - t_setup0;
  if (getenv("NSO_DEBUG")) fprintf(stderr, "[nso] total %.3f s, of which K_p CG %.3f s\n", t_total, t_kp);
  if (timings) { timings[0] = t_setup; timings[1] = t_total - t_setup; timings[2] = t_kp; timings[3] = t_total; }
  *iterations = it; *residual = res;
  free(V); free(w); free(tmp); free(tp); free(t2); free(y1); free(wk); free(H); free(cs); free(sn); free(g); free(yv);
  ilu_free(&iF); ilu_free(&iM); ilu_free(&iK);
  csr_free(&B);
  return out_of_time == 2 ? 3 : out_of_time ? 2 : rc;
}

/* M_p / K_p given on the FULL pattern (as assemble_impl writes them): extract the (1,1) blocks and solve. */
int nso_solve(int64_t N, int64_t n_u, const int64_t *rowptr, const int32_t *col, const double *A, const double *Mp,
              const double *Kp, const double *b, double nu, double rho, double deltat, double theta, int max_it,
              double tol_rel, int n_tmp_vectors, int nblocks, double schur_mass_coeff, double kp_tol, double *x, int *iterations,
              double *residual) {
  csr_t Mpp, Kpp;
  csr_block(rowptr, col, Mp, n_u, N, n_u, N, &Mpp);
  csr_block(rowptr, col, Kp, n_u, N, n_u, N, &Kpp);
  const int rc = nso_solve_blocks(N, n_u, rowptr, col, A, Mpp.ptr, Mpp.col, Mpp.val, Kpp.val, b, nu, rho, deltat, theta, max_it,
                                  tol_rel, n_tmp_vectors, nblocks, schur_mass_coeff, kp_tol, 0.0, x, iterations, residual, NULL);
  csr_free(&Mpp); csr_free(&Kpp);
  return rc;
}

void nso_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

int nso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* y = ILU(lof; nblocks row blocks)^-1 x for a CSR matrix: lets the tests check the factorization alone. */
void nso_ilu_apply_once(int n, const int64_t *ptr, const int32_t *col, const double *val, int lof, int nblocks, const double *x,
                        double *y, int64_t *nnz_factor) {
  ilu_t F;
  ilu_setup(&F, n, ptr, col, val, lof, nblocks);
  ilu_apply(&F, x, y);
  if (nnz_factor) *nnz_factor = F.ptr[n];
  ilu_free(&F);
}
