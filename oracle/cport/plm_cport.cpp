// Compiled CPU restatement of one SQP iteration of the reference (optimization/ocp.py:383-406):
//   sqp_data(x, p) -> OSQP update / solve -> Armijo line search.
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): bench.py times it as the CPU baseline ("port (C++)") and
// tests/test_cport.py pins it to the numpy oracle.  Nothing under pino_locoman_b200/ loads or links it.
//
//  * rows of g and the Jacobian values: the node mathematics of pino_locoman_b200/csrc/plm_node*.cuh compiled for the
//    host (the 32 lanes of a warp run in a loop per phase), one shooting node after the other, as casadi's VM walks
//    the nodes of sqp_data serially (optimization/ocp.py:386);
//  * OSQP (third-party, osqp 0.6.x semantics restated in SURVEY.md appendix A.8 and oracle/osqp_admm.py): Ruiz
//    equilibration on every update(Ax=), rho_vec classified with the row scaling in force at update_bounds, ADMM with
//    relaxation alpha, termination tests every check_termination iterations incl. the infeasibility certificates,
//    persistent scaled iterates.  The linear system is solved in the reduced form H = P + sigma I + A^T diag(rho) A
//    (block tridiagonal over the stages: only the integrator rows touch DX_{i+1}) with a dense block Cholesky -- the
//    same x~ as OSQP's QDLDL factorisation of the quasi-definite KKT matrix up to rounding, in fewer flops (so this
//    baseline is, if anything, faster than the real OSQP would be);
//  * Armijo line search with the reference's constants and its overwrite quirk (optimization/ocp.py:430-480).
//
// Build: oracle/cport/Makefile (g++ -O3 -march=x86-64-v3: AVX2 + FMA, runs on the build container and on the GPU box).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "../../pino_locoman_b200/csrc/plm_host.h"
#include "../../pino_locoman_b200/csrc/plm_node_driver.cuh"

using namespace plm;

namespace {

const double OSQP_INFTY = 1e30, MIN_SCALING = 1e-4, MAX_SCALING = 1e4, RHO_MIN = 1e-6, RHO_TOL = 1e-4, RHO_EQ = 1e3;

inline double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  return v > MAX_SCALING ? MAX_SCALING : v;
}

struct Port {
  HostTables t;
  plm_ocp_desc od;
  int n, m, nnz, N, ndx;
  // CSR / CSC of the fixed pattern of J_g (value order = CSR order of the product's pattern)
  std::vector<int> rptr, rcol, cptr, crow, csrc;
  // per stage: offset / size of the stage's variables, rows of its node
  std::vector<int> xo, so, ro, rn;
  // OSQP state
  std::vector<double> D, E, Eprev, P, q, Av, l, u, rho, x, z, y;
  double c = 1.0;
  // reduced factor: per stage Cholesky factor L_i (dense s x s, row major, lower) and W_i = L_i^-1 G_i^T (s x ndx)
  std::vector<std::vector<double>> Lf, Wf;
  std::vector<double> nodebuf;
  int iters = 0, status = 0;
  double t_eval = 0.0;
};

template <int KIND>
void run_node(Port& P, const double* x, const double* p, int node, double* g, double* Jv, int want_jac) {
  const HostTables& t = P.t;
  const PlmLayout& L = t.layout;
  const size_t need = node_ws_doubles(L, t.model.nv, L.nf, t.model.nbody, true) + 8;
  if (P.nodebuf.size() < need) P.nodebuf.assign(need, 0.0);
  NodeWs& ws = *reinterpret_cast<NodeWs*>(P.nodebuf.data());
  node_ws_bind(ws, L, t.model.nv, t.model.nbody, P.nodebuf.data() + (sizeof(NodeWs) + 7) / 8, nullptr);
  NodeArgs A;
  A.M = &t.model;
  A.L = &L;
  A.T = &L.types[L.node_type[node]];
  A.lut = t.lut.data() + A.T->lut_off;
  A.consts = t.consts.data() + A.T->const_off;
  A.xs = x + L.x_off[node];
  A.p = p;
  A.node = node;
  A.dt = node_dt(L, A.p, node);
  A.want_jac = want_jac;
  static thread_local HostExec ex;
  constexpr bool has_variant = KIND == PLM_CENTROIDAL_VEL || KIND == PLM_CENTROIDAL_ACC || KIND == PLM_WHOLE_BODY_ACC;
  if (has_variant && t.layout.nobase) node_eval_body<KIND, has_variant>(ex, ws, A);      // include_base = False
  else node_eval_body<KIND, false>(ex, ws, A);
  const PlmNodeType& T = *A.T;
  for (int r = 0; r < T.nrows; ++r) g[L.row_off[node] + r] = ws.g[r];
  if (node == 0) for (int r = 0; r < L.ndx; ++r) g[r] = A.xs[r];
  if (want_jac) {
    for (int e = 0; e < T.nnz; ++e) Jv[L.nnz_off[node] + e] = ws.J[e];
    if (node == 0) for (int e = 0; e < L.ndx; ++e) Jv[e] = 1.0;
  }
}

void eval_nodes(Port& P, const double* x, const double* p, double* g, double* Jv, int want_jac) {
  auto t0 = std::chrono::steady_clock::now();
  const PlmLayout& L = P.t.layout;
  for (int i = 0; i < L.nodes; ++i) switch (L.dynamics) {
      case PLM_CENTROIDAL_VEL: run_node<PLM_CENTROIDAL_VEL>(P, x, p, i, g, Jv, want_jac); break;
      case PLM_CENTROIDAL_ACC: run_node<PLM_CENTROIDAL_ACC>(P, x, p, i, g, Jv, want_jac); break;
      case PLM_WHOLE_BODY_ACC: run_node<PLM_WHOLE_BODY_ACC>(P, x, p, i, g, Jv, want_jac); break;
      case PLM_WHOLE_BODY_ABA: run_node<PLM_WHOLE_BODY_ABA>(P, x, p, i, g, Jv, want_jac); break;
      default: run_node<PLM_WHOLE_BODY_RNEA>(P, x, p, i, g, Jv, want_jac); break;
    }
  P.t_eval += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- objective: f = sum w1 (x - t1)^2 + sum w2 (x - t2)^2 (second term: the previous-torque cost of whole_body_rnea)
double objective(const Port& P, const double* x, const double* w1, const double* t1, const double* w2, const double* t2, double* grad) {
  double f = 0.0;
  for (int j = 0; j < P.n; ++j) {
    const double e1 = x[j] - t1[j], e2 = x[j] - t2[j];
    f += w1[j] * e1 * e1 + w2[j] * e2 * e2;
    if (grad) grad[j] = 2.0 * w1[j] * e1 + 2.0 * w2[j] * e2;
  }
  return f;
}

// ---- OSQP: scale_data() of a fresh update(Ax=) (oracle/osqp_admm.py:_scale_data)
void scale_data(Port& S, const double* P0, const double* q0, const double* A0, const double* l0, const double* u0, bool dummy) {
  const int n = S.n, m = S.m;
  std::vector<double>&D = S.D, &E = S.E;
  std::fill(D.begin(), D.end(), 1.0);
  std::fill(E.begin(), E.end(), 1.0);
  double c = 1.0;
  std::vector<double> Dn(n), En(m);
  for (int pass = 0; pass < S.od.osqp_scaling; ++pass) {
    for (int r = 0; r < m; ++r) {
      double v = 0.0;
      for (int e = S.rptr[r]; e < S.rptr[r + 1]; ++e) v = std::max(v, (dummy ? 1.0 : fabs(A0[e])) * D[S.rcol[e]]);
      En[r] = E[r] / sqrt(limit_scaling(E[r] * v));
    }
    for (int j = 0; j < n; ++j) {
      double v = 0.0;
      for (int e = S.cptr[j]; e < S.cptr[j + 1]; ++e) v = std::max(v, (dummy ? 1.0 : fabs(A0[S.csrc[e]])) * E[S.crow[e]]);
      v = std::max(D[j] * v, c * D[j] * D[j] * fabs(P0[j]));
      Dn[j] = D[j] / sqrt(limit_scaling(v));
    }
    D = Dn;
    E = En;
    double sp = 0.0, mq = 0.0;
    for (int j = 0; j < n; ++j) {
      sp += c * D[j] * D[j] * fabs(P0[j]);
      mq = std::max(mq, fabs(c * D[j] * (dummy ? 1.0 : q0[j])));
    }
    const double ct = 1.0 / limit_scaling(std::max(sp / n, limit_scaling(mq)));
    c *= ct;
  }
  S.c = c;
  if (dummy) return;
  for (int j = 0; j < n; ++j) { S.P[j] = c * D[j] * D[j] * P0[j]; S.q[j] = c * D[j] * q0[j]; }
  for (int r = 0; r < m; ++r)
    for (int e = S.rptr[r]; e < S.rptr[r + 1]; ++e) S.Av[e] = E[r] * A0[e] * D[S.rcol[e]];
  for (int r = 0; r < m; ++r) { S.l[r] = E[r] * l0[r]; S.u[r] = E[r] * u0[r]; }
}

// ---- reduced block-tridiagonal factorisation: H_ii = P_i + sigma + A_i^T R A_i (+ coupling of the previous node's
// integrator rows), H_{i+1,i} from rows touching DX_{i+1}.  Dense per stage.
bool factor(Port& S) {
  const int N = S.N, ndx = S.ndx;
  const double sigma = S.od.osqp_sigma;
  std::vector<double> Hn;            // (s + ndx) x (s + ndx) local normal matrix of node i
  std::vector<double> carry(ndx * ndx, 0.0), K(ndx * ndx, 0.0);
  S.Lf.assign(N + 1, {});
  S.Wf.assign(N + 1, {});
  // rows 0..ndx-1 (DX_0 == 0): one entry each on DX_0
  std::vector<double> d0(ndx, 0.0);
  for (int r = 0; r < ndx; ++r) { const double a = S.Av[S.rptr[r]]; d0[S.rcol[S.rptr[r]]] += S.rho[r] * a * a; }
  for (int i = 0; i <= N; ++i) {
    const int s = S.so[i], xo = S.xo[i];
    const int w = (i < N) ? s + ndx : s;
    Hn.assign((size_t)w * w, 0.0);
    for (int j = 0; j < s; ++j) Hn[(size_t)j * w + j] = S.P[xo + j] + sigma;
    if (i == 0) for (int j = 0; j < ndx; ++j) Hn[(size_t)j * w + j] += d0[j];
    else
      for (int a = 0; a < ndx; ++a)
        for (int b = 0; b < ndx; ++b) Hn[(size_t)a * w + b] += carry[a * ndx + b] - K[a * ndx + b];
    if (i < N) {
      for (int r = S.ro[i]; r < S.ro[i] + S.rn[i]; ++r) {
        const double rr = S.rho[r];
        for (int e1 = S.rptr[r]; e1 < S.rptr[r + 1]; ++e1) {
          const int a = S.rcol[e1] - xo;
          const double wa = rr * S.Av[e1];
          for (int e2 = S.rptr[r]; e2 < S.rptr[r + 1]; ++e2) Hn[(size_t)a * w + (S.rcol[e2] - xo)] += wa * S.Av[e2];
        }
      }
    }
    // Cholesky of the s x s leading block
    std::vector<double>& Lc = S.Lf[i];
    Lc.assign((size_t)s * s, 0.0);
    for (int j = 0; j < s; ++j) {
      double d = Hn[(size_t)j * w + j];
      for (int k = 0; k < j; ++k) d -= Lc[(size_t)j * s + k] * Lc[(size_t)j * s + k];
      if (!(d > 0.0)) return false;
      const double dj = sqrt(d);
      Lc[(size_t)j * s + j] = dj;
      for (int r = j + 1; r < s; ++r) {
        double v = Hn[(size_t)r * w + j];
        const double* lr = &Lc[(size_t)r * s];
        const double* lj = &Lc[(size_t)j * s];
        for (int k = 0; k < j; ++k) v -= lr[k] * lj[k];
        Lc[(size_t)r * s + j] = v / dj;
      }
    }
    if (i == N) break;
    // W = L^-1 G^T (G = H_{i+1,i}: ndx x s), K = W^T W, carry = H_{i+1,i+1} contribution of node i
    std::vector<double>& W = S.Wf[i];
    W.assign((size_t)s * ndx, 0.0);
    for (int cidx = 0; cidx < ndx; ++cidx) {
      for (int j = 0; j < s; ++j) {
        double v = Hn[(size_t)(s + cidx) * w + j];
        const double* lj = &Lc[(size_t)j * s];
        for (int k = 0; k < j; ++k) v -= lj[k] * W[(size_t)k * ndx + cidx];
        W[(size_t)j * ndx + cidx] = v / lj[j];
      }
    }
    for (int a = 0; a < ndx; ++a)
      for (int b = 0; b < ndx; ++b) {
        double v = 0.0;
        for (int k = 0; k < s; ++k) v += W[(size_t)k * ndx + a] * W[(size_t)k * ndx + b];
        K[a * ndx + b] = v;
        carry[a * ndx + b] = Hn[(size_t)(s + a) * w + (s + b)];
      }
  }
  return true;
}

// x~ = H^-1 rhs (in place)
void solve_reduced(const Port& S, double* v) {
  const int N = S.N, ndx = S.ndx;
  for (int i = 0; i <= N; ++i) {          // forward: y_i = L_i^-1 (b_i - W_{i-1}^T y_{i-1})
    const int s = S.so[i];
    double* b = v + S.xo[i];
    if (i > 0) {
      const int sp = S.so[i - 1];
      const double* yp = v + S.xo[i - 1];
      const std::vector<double>& W = S.Wf[i - 1];
      for (int k = 0; k < sp; ++k) {
        const double yk = yp[k];
        const double* wk = &W[(size_t)k * ndx];
        for (int a = 0; a < ndx; ++a) b[a] -= wk[a] * yk;
      }
    }
    const std::vector<double>& Lc = S.Lf[i];
    for (int j = 0; j < s; ++j) {
      double t = b[j];
      const double* lj = &Lc[(size_t)j * s];
      for (int k = 0; k < j; ++k) t -= lj[k] * b[k];
      b[j] = t / lj[j];
    }
  }
  for (int i = N; i >= 0; --i) {          // backward: x_i = L_i^-T (y_i - W_i x_{i+1}[0:ndx])
    const int s = S.so[i];
    double* b = v + S.xo[i];
    if (i < N) {
      const double* xn = v + S.xo[i + 1];
      const std::vector<double>& W = S.Wf[i];
      for (int k = 0; k < s; ++k) {
        const double* wk = &W[(size_t)k * ndx];
        double t = 0.0;
        for (int a = 0; a < ndx; ++a) t += wk[a] * xn[a];
        b[k] -= t;
      }
    }
    const std::vector<double>& Lc = S.Lf[i];
    for (int j = s - 1; j >= 0; --j) {
      const double t = b[j] / Lc[(size_t)j * s + j];
      b[j] = t;
      const double* lj = &Lc[(size_t)j * s];
      for (int k = 0; k < j; ++k) b[k] -= lj[k] * t;
    }
  }
}

void spmv(const Port& S, const double* v, double* out) {          // out = A v
  for (int r = 0; r < S.m; ++r) {
    double a = 0.0;
    for (int e = S.rptr[r]; e < S.rptr[r + 1]; ++e) a += S.Av[e] * v[S.rcol[e]];
    out[r] = a;
  }
}
void spmv_t(const Port& S, const double* w, double* out) {        // out = A^T w
  for (int j = 0; j < S.n; ++j) {
    double a = 0.0;
    for (int e = S.cptr[j]; e < S.cptr[j + 1]; ++e) a += S.Av[S.csrc[e]] * w[S.crow[e]];
    out[j] = a;
  }
}
double ninf(const std::vector<double>& v) { double a = 0.0; for (double e : v) a = std::max(a, fabs(e)); return a; }

// status: 1 solved, 2 solved inaccurate, -3/3 primal infeasible (/inaccurate), -4/4 dual infeasible, 0 unsolved
int check_termination(const Port& S, const std::vector<double>& x, const std::vector<double>& z, const std::vector<double>& y,
                      const std::vector<double>& dxv, const std::vector<double>& dyv, bool approx) {
  const int n = S.n, m = S.m;
  const double k = approx ? 10.0 : 1.0;
  const double eps_abs = S.od.osqp_eps_abs * k, eps_rel = S.od.osqp_eps_rel * k, eps_pinf = S.od.osqp_eps_prim_inf * k,
               eps_dinf = S.od.osqp_eps_dual_inf * k;
  std::vector<double> Ax(m), Aty(n);
  spmv(S, x.data(), Ax.data());
  spmv_t(S, y.data(), Aty.data());
  double pri = 0, nz = 0, nax = 0;
  for (int r = 0; r < m; ++r) {
    pri = std::max(pri, fabs((Ax[r] - z[r]) / S.E[r]));
    nz = std::max(nz, fabs(z[r] / S.E[r]));
    nax = std::max(nax, fabs(Ax[r] / S.E[r]));
  }
  double dua = 0, nq = 0, naty = 0, npx = 0;
  for (int j = 0; j < n; ++j) {
    const double px = S.P[j] * x[j];
    dua = std::max(dua, fabs((px + S.q[j] + Aty[j]) / S.D[j]));
    nq = std::max(nq, fabs(S.q[j] / S.D[j]));
    naty = std::max(naty, fabs(Aty[j] / S.D[j]));
    npx = std::max(npx, fabs(px / S.D[j]));
  }
  dua /= S.c;
  const double eps_pri = eps_abs + eps_rel * std::max(nz, nax);
  const double eps_dua = eps_abs + eps_rel * std::max(nq, std::max(naty, npx)) / S.c;
  const bool prim_ok = pri < eps_pri, dual_ok = dua < eps_dua;
  bool prim_inf = false, dual_inf = false;
  if (!prim_ok) {
    std::vector<double> dy(dyv);
    double norm = 0.0, lhs = 0.0;
    for (int r = 0; r < m; ++r) {
      const bool up = S.u[r] > OSQP_INFTY * MIN_SCALING, lo = S.l[r] < -OSQP_INFTY * MIN_SCALING;
      if (up && lo) dy[r] = 0.0;
      else if (up) dy[r] = std::min(dy[r], 0.0);
      else if (lo) dy[r] = std::max(dy[r], 0.0);
      norm = std::max(norm, fabs(S.E[r] * dy[r]));
    }
    if (norm > eps_pinf) {
      for (int r = 0; r < m; ++r) lhs += S.u[r] * std::max(dy[r], 0.0) + S.l[r] * std::min(dy[r], 0.0);
      if (lhs < -eps_pinf * norm) {
        std::vector<double> at(n);
        spmv_t(S, dy.data(), at.data());
        double na = 0.0;
        for (int j = 0; j < n; ++j) na = std::max(na, fabs(at[j] / S.D[j]));
        prim_inf = na < eps_pinf * norm;
      }
    }
  }
  if (!dual_ok) {
    double norm = 0.0;
    for (int j = 0; j < n; ++j) norm = std::max(norm, fabs(S.D[j] * dxv[j]));
    if (norm > eps_dinf) {
      double qd = 0.0, npd = 0.0;
      for (int j = 0; j < n; ++j) { qd += S.q[j] * dxv[j]; npd = std::max(npd, fabs(S.P[j] * dxv[j] / S.D[j])); }
      if (qd < -S.c * eps_dinf * norm && npd < S.c * eps_dinf * norm) {
        std::vector<double> adx(m);
        spmv(S, dxv.data(), adx.data());
        bool bad = false;
        for (int r = 0; r < m; ++r) {
          const double a = adx[r] / S.E[r];
          if ((S.u[r] < OSQP_INFTY * MIN_SCALING && a > eps_dinf * norm) || (S.l[r] > -OSQP_INFTY * MIN_SCALING && a < -eps_dinf * norm)) bad = true;
        }
        dual_inf = !bad;
      }
    }
  }
  if (prim_ok && dual_ok) return approx ? 2 : 1;
  if (prim_inf) return approx ? 3 : -3;
  if (dual_inf) return approx ? 4 : -4;
  return 0;
}

// osqp.update(q=, l=, u=, Ax=) in the python wrapper's order, then solve(); returns the unscaled step in dx
void osqp_update_solve(Port& S, const double* P0, const double* q0, const double* A0, const double* l0in, const double* u0in, double* dx) {
  const int n = S.n, m = S.m;
  std::vector<double> l0(m), u0(m);
  for (int r = 0; r < m; ++r) { l0[r] = std::max(l0in[r], -OSQP_INFTY); u0[r] = std::min(u0in[r], OSQP_INFTY); }
  // update_bounds: rows classified with the row scaling of the previous data
  for (int r = 0; r < m; ++r) {
    const double lp = S.Eprev[r] * l0[r], up = S.Eprev[r] * u0[r];
    if (lp < -OSQP_INFTY * MIN_SCALING && up > OSQP_INFTY * MIN_SCALING) S.rho[r] = RHO_MIN;
    else if (up - lp < RHO_TOL) S.rho[r] = RHO_EQ * S.od.osqp_rho;
    else S.rho[r] = S.od.osqp_rho;
  }
  scale_data(S, P0, q0, A0, l0.data(), u0.data(), false);
  S.Eprev = S.E;
  const bool ok = factor(S);
  const double alpha = S.od.osqp_alpha, sigma = S.od.osqp_sigma;
  std::vector<double>&x = S.x, &z = S.z, &y = S.y;
  std::vector<double> xt(n), w(m), zt(m), dxv(n), dyv(m);
  int status = 0, it = 0;
  for (it = 1; it <= S.od.osqp_max_iter; ++it) {
    for (int r = 0; r < m; ++r) w[r] = S.rho[r] * z[r] - y[r];
    spmv_t(S, w.data(), xt.data());
    for (int j = 0; j < n; ++j) xt[j] += sigma * x[j] - S.q[j];
    solve_reduced(S, xt.data());
    spmv(S, xt.data(), zt.data());
    for (int j = 0; j < n; ++j) {
      const double xn = alpha * xt[j] + (1.0 - alpha) * x[j];
      dxv[j] = xn - x[j];
      x[j] = xn;
    }
    for (int r = 0; r < m; ++r) {
      const double zr = alpha * zt[r] + (1.0 - alpha) * z[r];
      double zn = zr + y[r] / S.rho[r];
      zn = std::min(std::max(zn, S.l[r]), S.u[r]);
      dyv[r] = S.rho[r] * (zr - zn);
      y[r] += dyv[r];
      z[r] = zn;
    }
    if (S.od.osqp_check_termination > 0 && it % S.od.osqp_check_termination == 0) {
      status = check_termination(S, x, z, y, dxv, dyv, false);
      if (status != 0) break;
    }
  }
  if (it > S.od.osqp_max_iter) it = S.od.osqp_max_iter;
  if (status == 0) {
    status = check_termination(S, x, z, y, dxv, dyv, true);
    if (status == 0) status = -2;
  }
  if (!ok) status = -10;
  S.iters = it;
  S.status = status;
  const bool none = status == 3 || status == -3 || status == 4 || status == -4;
  for (int j = 0; j < n; ++j) dx[j] = none ? nan("") : S.D[j] * x[j];
  if (none) { std::fill(x.begin(), x.end(), 0.0); std::fill(z.begin(), z.end(), 0.0); std::fill(y.begin(), y.end(), 0.0); }
}

void violation(const Port& S, const double* g, const double* lbg, const double* ubg, double* metric, double* vmax) {
  double ss = 0.0, mx = 0.0;
  for (int r = 0; r < S.m; ++r) {
    const double v1 = std::max(0.0, lbg[r] - g[r]), v2 = std::max(0.0, g[r] - ubg[r]);
    ss += v1 * v1 + v2 * v2;
    mx = std::max(mx, std::max(v1, v2));
  }
  *metric = sqrt(ss);
  *vmax = mx;
}

}  // namespace

extern "C" {

void* cport_create(const plm_robot_desc* r, const plm_ocp_desc* o, char* err, int errlen) {
  Port* P = new Port();
  P->od = *o;
  if (!build_tables(*r, *o, P->t)) {
    strncpy(err, P->t.error.c_str(), errlen - 1);
    delete P;
    return nullptr;
  }
  const PlmLayout& L = P->t.layout;
  P->n = L.n; P->m = L.m; P->nnz = L.nnz; P->N = L.nodes; P->ndx = L.ndx;
  const int n = P->n, m = P->m, nnz = P->nnz;
  // pattern (value order is row-major CSR)
  P->rptr.assign(m + 1, 0);
  P->rcol.resize(nnz);
  for (int e = 0; e < nnz; ++e) { P->rptr[P->t.pat_rows[e] + 1]++; P->rcol[e] = P->t.pat_cols[e]; }
  for (int r = 0; r < m; ++r) P->rptr[r + 1] += P->rptr[r];
  for (int e = 1; e < nnz; ++e)
    if (P->t.pat_rows[e] < P->t.pat_rows[e - 1]) { strncpy(err, "pattern is not row-major", errlen - 1); delete P; return nullptr; }
  P->cptr.assign(n + 1, 0);
  for (int e = 0; e < nnz; ++e) P->cptr[P->rcol[e] + 1]++;
  for (int j = 0; j < n; ++j) P->cptr[j + 1] += P->cptr[j];
  P->crow.resize(nnz);
  P->csrc.resize(nnz);
  {
    std::vector<int> fill(P->cptr.begin(), P->cptr.end() - 1);
    for (int e = 0; e < nnz; ++e) { const int j = P->rcol[e], k = fill[j]++; P->crow[k] = P->t.pat_rows[e]; P->csrc[k] = e; }
  }
  for (int i = 0; i <= L.nodes; ++i) {
    P->xo.push_back(L.x_off[i]);
    P->so.push_back(i < L.nodes ? L.x_off[i + 1] - L.x_off[i] : L.ndx);
    P->ro.push_back(i < L.nodes ? L.row_off[i] : 0);
    P->rn.push_back(i < L.nodes ? L.types[L.node_type[i]].nrows : 0);
  }
  P->D.assign(n, 1.0); P->E.assign(m, 1.0); P->Eprev.assign(m, 1.0); P->P.assign(n, 0.0); P->q.assign(n, 0.0);
  P->Av.assign(nnz, 0.0); P->l.assign(m, 0.0); P->u.assign(m, 0.0); P->rho.assign(m, 0.0);
  P->x.assign(n, 0.0); P->z.assign(m, 0.0); P->y.assign(m, 0.0);
  return P;
}

void cport_destroy(void* h) { delete static_cast<Port*>(h); }

void cport_dims(void* h, int* out) {
  Port* P = static_cast<Port*>(h);
  const PlmLayout& L = P->t.layout;
  int v[6] = {L.n, L.m, L.np, L.nnz, L.ndx, L.nodes};
  memcpy(out, v, sizeof(v));
}

void cport_pattern(void* h, int* rows, int* cols) {
  Port* P = static_cast<Port*>(h);
  memcpy(rows, P->t.pat_rows.data(), P->nnz * sizeof(int));
  memcpy(cols, P->t.pat_cols.data(), P->nnz * sizeof(int));
}

// osqp setup() with the dummy data of optimization/ocp.py:305-313: zero iterates, setup-time row scaling
void cport_init_solver(void* h, const double* hess) {
  Port* P = static_cast<Port*>(h);
  std::fill(P->x.begin(), P->x.end(), 0.0);
  std::fill(P->z.begin(), P->z.end(), 0.0);
  std::fill(P->y.begin(), P->y.end(), 0.0);
  scale_data(*P, hess, nullptr, nullptr, nullptr, nullptr, true);
  P->Eprev = P->E;
}

void cport_node_eval(void* h, const double* x, const double* p, double* g, double* Jv, int want_jac) {
  eval_nodes(*static_cast<Port*>(h), x, p, g, Jv, want_jac);
}

// One SQP iteration.  hess [n] = diag of the objective Hessian; w1/t1/w2/t2 [n] the objective (see objective());
// lbg/ubg [m].  Outputs: x_new [n], dx [n] (QP step), info[8] = {ADMM iterations, OSQP status, accepted, step size,
// trials, f, g_metric, violation_max}.
void cport_sqp_iteration(void* h, const double* x, const double* p, const double* hess, const double* w1, const double* t1,
                         const double* w2, const double* t2, const double* lbg, const double* ubg, double* x_new, double* dx, double* info) {
  Port& S = *static_cast<Port*>(h);
  const int n = S.n, m = S.m;
  std::vector<double> g(m), Jv(S.nnz), grad(n), l(m), u(m);
  eval_nodes(S, x, p, g.data(), Jv.data(), 1);
  objective(S, x, w1, t1, w2, t2, grad.data());
  for (int r = 0; r < m; ++r) { l[r] = lbg[r] - g[r]; u[r] = ubg[r] - g[r]; }
  osqp_update_solve(S, hess, grad.data(), Jv.data(), l.data(), u.data(), dx);
  // Armijo line search (optimization/ocp.py:430-480)
  const double armijo_factor = 1e-4, a_min = 1e-4, a_decay = 0.5, g_max = 1e-3, g_min = 1e-5, gamma = 1e-5;
  double a = 1.0;
  double f = objective(S, x, w1, t1, w2, t2, nullptr);
  double g_metric, vmax;
  violation(S, g.data(), lbg, ubg, &g_metric, &vmax);
  double armijo_metric = 0.0;
  for (int j = 0; j < n; ++j) armijo_metric += grad[j] * dx[j];
  bool accepted = false;
  int trials = 0;
  std::vector<double> xn(x, x + n), gn(m);
  while (!accepted && a > a_min) {
    for (int j = 0; j < n; ++j) xn[j] = x[j] + a * dx[j];
    const double new_f = objective(S, xn.data(), w1, t1, w2, t2, nullptr);
    eval_nodes(S, xn.data(), p, gn.data(), nullptr, 0);
    ++trials;
    double new_metric, nm;
    violation(S, gn.data(), lbg, ubg, &new_metric, &nm);
    if (new_metric > g_max) {
      if (new_metric < (1 - gamma) * g_metric) accepted = true;
    } else if (std::max(new_metric, g_metric) < g_min && armijo_metric < 0) {
      if (new_f <= f + armijo_factor * armijo_metric) accepted = true;
    } else if (new_f <= f - gamma * new_metric || new_metric < (1 - gamma) * g_metric) accepted = true;
    a *= a_decay;
    f = new_f;
    g_metric = new_metric;
  }
  if (accepted) memcpy(x_new, xn.data(), n * sizeof(double));
  else memcpy(x_new, x, n * sizeof(double));
  eval_nodes(S, x_new, p, gn.data(), nullptr, 0);      // the violation print of optimization/ocp.py:412-414
  double fm, fv;
  violation(S, gn.data(), lbg, ubg, &fm, &fv);
  info[0] = S.iters; info[1] = S.status; info[2] = accepted ? 1.0 : 0.0; info[3] = a / a_decay; info[4] = trials;
  info[5] = f; info[6] = g_metric; info[7] = fv;
}

double cport_eval_seconds(void* h, int reset) {
  Port* P = static_cast<Port*>(h);
  const double t = P->t_eval;
  if (reset) P->t_eval = 0.0;
  return t;
}

void cport_get_scaling(void* h, double* D, double* E, double* c) {
  Port* P = static_cast<Port*>(h);
  memcpy(D, P->D.data(), P->n * sizeof(double));
  memcpy(E, P->E.data(), P->m * sizeof(double));
  *c = P->c;
}

void cport_get_iterates(void* h, double* x, double* z, double* y) {
  Port* P = static_cast<Port*>(h);
  memcpy(x, P->x.data(), P->n * sizeof(double));
  memcpy(z, P->z.data(), P->m * sizeof(double));
  memcpy(y, P->y.data(), P->m * sizeof(double));
}

}  // extern "C"
