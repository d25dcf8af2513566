// CUDA kernels of the evaluation side of the SQP iteration: node rows + Jacobian blocks, bounds, objective.
#pragma once
#include <cuda_runtime.h>

#include "plm_node_driver.cuh"

namespace plm {

// Line-search trial mode of the node kernel: evaluate rows at x + alphas[t] * dx for t in [t0, t0 + ntrial) and
// reduce the constraint violation per (instance, trial, node) instead of writing g.
struct TrialArgs {
  const double* dxs;      // [batch][n] search direction
  const double* alphas;   // [ntot] step sizes
  int t0, ntrial, ntot;
  const int* accepted;    // [batch] skip instances whose line search has finished (may be null)
  const double* lbg;      // [batch][m]
  const double* ubg;
  double* part;           // [batch][ntot][nodes][2] (sum of squares, max); null => normal mode
};

// One warp per (instance, node).  Shared memory: [PlmModel | PlmLayout | per-warp NodeWs + row/J staging].
template <int KIND, bool NOBASE>
__global__ void __launch_bounds__(PLM_NODE_WARPS * 32, 2)
node_eval_kernel(DeviceTables tab, const double* __restrict__ x, const double* __restrict__ p, int batch,
                 double* __restrict__ g, double* __restrict__ Jv, int want_jac, int ws_doubles, TrialArgs tr) {
  extern __shared__ double smem[];
  PlmModel* sM = reinterpret_cast<PlmModel*>(smem);
  PlmLayout* sL = reinterpret_cast<PlmLayout*>(smem + (sizeof(PlmModel) + 7) / 8);
  double* wsbase = smem + (sizeof(PlmModel) + 7) / 8 + (sizeof(PlmLayout) + 7) / 8;
  if (tr.part && tr.accepted) {
    // later line-search launches: a CTA whose instances have all finished leaves before staging the tables
    const PlmLayout& Lg = *tab.layout;
    const long long it0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool live = it0 < (long long)batch * tr.ntrial * Lg.nodes && !tr.accepted[(int)(it0 / Lg.nodes) / tr.ntrial];
    if (!__syncthreads_or(live)) return;
  }
  {
    const int* src = reinterpret_cast<const int*>(tab.model);
    int* dst = reinterpret_cast<int*>(sM);
    for (int i = threadIdx.x; i < (int)(sizeof(PlmModel) / 4); i += blockDim.x) dst[i] = src[i];
    src = reinterpret_cast<const int*>(tab.layout);
    dst = reinterpret_cast<int*>(sL);
    for (int i = threadIdx.x; i < (int)(sizeof(PlmLayout) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PlmLayout& L = *sL;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int ntr = tr.part ? tr.ntrial : 1;
  if (item >= (long long)batch * ntr * L.nodes) return;
  const int node = (int)(item % L.nodes);
  const int bt = (int)(item / L.nodes);
  const int b = bt / ntr, trial = tr.t0 + bt % ntr;
  if (tr.part && tr.accepted && tr.accepted[b]) return;   // line search already finished for this instance
  NodeWs& ws = *reinterpret_cast<NodeWs*>(wsbase + (size_t)warp * ws_doubles);
  // Jacobian entries go straight to the instance's node block in HBM (every pattern entry is written exactly once)
  if (lane == 0)
    node_ws_bind(ws, L, sM->nv, sM->nbody, wsbase + (size_t)warp * ws_doubles + (sizeof(NodeWs) + 7) / 8,
                 want_jac ? Jv + (size_t)b * L.nnz + L.nnz_off[node] : wsbase /* unused */, tr.part != nullptr);
  __syncwarp();
  NodeArgs A;
  A.M = sM;
  A.L = sL;
  A.T = &L.types[L.node_type[node]];
  A.lut = tab.lut + A.T->lut_off;
  A.consts = tab.consts + A.T->const_off;
  A.xs = x + (size_t)b * L.n + L.x_off[node];
  if (tr.part) {   // line-search trial: stage x + alpha dx of this node in shared memory
    const double alpha = tr.alphas[trial];
    const double* dxp = tr.dxs + (size_t)b * L.n + L.x_off[node];
    const int len = 2 * L.ndx + L.types[L.node_type[node]].nu;
    for (int k2 = lane; k2 < len; k2 += 32) ws.xbuf[k2] = A.xs[k2] + alpha * dxp[k2];
    __syncwarp();
    A.xs = ws.xbuf;
  }
  A.p = p + (size_t)b * L.np;
  A.node = node;
  A.dt = node_dt(L, A.p, node);
  A.want_jac = want_jac;
  WarpExec ex;
  ex.lane = lane;
  node_eval_body<KIND, NOBASE>(ex, ws, A);
  const PlmNodeType& T = *A.T;
  if (tr.part) {
    // constraint-violation partials of this node: sum of squares and max of [max(0, lbg-g); max(0, g-ubg)]
    const double* lo = tr.lbg + (size_t)b * L.m + L.row_off[node];
    const double* up = tr.ubg + (size_t)b * L.m + L.row_off[node];
    double ss = 0.0, mx = 0.0;
    for (int r = lane; r < T.nrows; r += 32) {
      const double gv = ws.g[r];
      const double v1 = fmax(0.0, lo[r] - gv), v2 = fmax(0.0, gv - up[r]);
      ss += v1 * v1 + v2 * v2;
      mx = fmax(mx, fmax(v1, v2));
    }
    if (node == 0) {
      for (int r = lane; r < L.ndx; r += 32) {   // DX_0 == 0 rows
        const double gv = A.xs[r];
        ss += gv * gv;
        mx = fmax(mx, fabs(gv));
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_down_sync(0xffffffffu, ss, o);
      mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      double* po = tr.part + 2 * (((size_t)b * tr.ntot + trial) * L.nodes + node);
      po[0] = ss;
      po[1] = mx;
    }
    return;
  }
  // coalesced write-out
  double* go = g + (size_t)b * L.m + L.row_off[node];
  for (int r = lane; r < T.nrows; r += 32) store_j(go + r, ws.g[r]);
  if (node == 0) {   // DX_0 == 0 rows (optimization/ocp.py:109)
    double* g0 = g + (size_t)b * L.m;
    for (int r = lane; r < L.ndx; r += 32) store_j(g0 + r, A.xs[r]);
  }
  if (want_jac) {
    if (node == 0) {
      double* J0 = Jv + (size_t)b * L.nnz;
      for (int e = lane; e < L.ndx; e += 32) store_j(J0 + e, 1.0);
    }
  }
}

// lbg / ubg: functions of p only.  One warp per (instance, node); node 0 also writes the DX_0 rows.
__global__ void bounds_kernel(DeviceTables tab, const double* __restrict__ p, int batch,
                              double* __restrict__ lbg, double* __restrict__ ubg) {
  const PlmModel& M = *tab.model;
  const PlmLayout& L = *tab.layout;
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= (long long)batch * L.nodes) return;
  const int b = (int)(item / L.nodes), node = (int)(item % L.nodes);
  const PlmNodeType& T = L.types[L.node_type[node]];
  const double* pp = p + (size_t)b * L.np;
  double* lo = lbg + (size_t)b * L.m + L.row_off[node];
  double* up = ubg + (size_t)b * L.m + L.row_off[node];
  const double inf = INFINITY;
  for (int r = lane; r < T.nrows; r += 32) {
    double l = 0.0, u = 0.0;
    if (T.row_taub >= 0 && r >= T.row_taub && r < T.row_taub + M.nj) {
      u = M.joint_torque_max[r - T.row_taub]; l = -u;
    } else if (r >= T.row_foot[0] && r < T.row_foot[3] + (T.state_rows ? 8 : 5)) {
      const int per = T.state_rows ? 8 : 5;
      const int lr = (r - T.row_foot[0]) % per;
      if (lr == 0) u = inf;
      else if (lr == 1) l = -inf;
    } else if (T.row_ext >= 0 && r >= T.row_ext && r < T.row_ext + 3) {
      l = u = pp[L.p_ext_force + r - T.row_ext];
    } else if (T.row_qj >= 0 && r >= T.row_qj && r < T.row_qj + M.nj) {
      l = M.joint_pos_min[r - T.row_qj]; u = M.joint_pos_max[r - T.row_qj];
    } else if (T.row_vj >= 0 && r >= T.row_vj && r < T.row_vj + M.nj) {
      u = M.joint_vel_max[r - T.row_vj]; l = -u;
    }
    lo[r] = l;
    up[r] = u;
  }
  if (node == 0) {
    for (int r = lane; r < L.ndx; r += 32) {
      lbg[(size_t)b * L.m + r] = 0.0;
      ubg[(size_t)b * L.m + r] = 0.0;
    }
  }
}

// Per-instance tracking targets (setup_targets of each ocp_*.py): tgt[b][0:ndx] = dx_des, tgt[b][ndx:ndx+nu0] = u_des
__global__ void targets_kernel(DeviceTables tab, const double* __restrict__ p, int batch, double* __restrict__ tgt, int tgt_ld) {
  const PlmModel& M = *tab.model;
  const PlmLayout& L = *tab.layout;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* pp = p + (size_t)b * L.np;
  const double* x_init = pp + L.p_x_init;
  double* t = tgt + (size_t)b * tgt_ld;
  const int nv = M.nv, nj = M.nj;
  const bool cvel = L.dynamics == PLM_CENTROIDAL_VEL;
  const int qoff = cvel ? 6 : 0;
  double nu6[6];
  se3_difference(x_init + qoff, x_init + qoff + 3, M.q0, M.q0 + 3, nu6);
  if (cvel) {
    for (int i = 0; i < 6; ++i) t[i] = pp[L.p_base_vel + i] - x_init[i];
    for (int i = 0; i < 6; ++i) t[6 + i] = nu6[i];
    for (int j = 0; j < nj; ++j) t[12 + j] = M.q0[7 + j] - x_init[6 + 7 + j];
  } else {
    for (int i = 0; i < 6; ++i) t[i] = nu6[i];
    for (int j = 0; j < nj; ++j) t[6 + j] = M.q0[7 + j] - x_init[7 + j];
    for (int i = 0; i < 6; ++i) t[nv + i] = pp[L.p_base_vel + i] - x_init[M.nq + i];
    for (int j = 0; j < nj; ++j) t[nv + 6 + j] = 0.0 - x_init[M.nq + 6 + j];
  }
  double* ud = t + L.ndx;
  const int nu0 = L.types[L.node_type[0]].nu;
  for (int i = 0; i < nu0; ++i) ud[i] = 0.0;
  const double fg = 9.81 * M.total_mass;
  const double nc = pp[L.p_n_contacts];
  ud[L.f_idx + 2] = ud[L.f_idx + 5] = 0.8 * fg / nc;
  ud[L.f_idx + 8] = ud[L.f_idx + 11] = 1.2 * fg / nc;
}

// Weight / target of decision variable k of x for the instance (objective of setup_objective).
__device__ inline void var_weight_target(const PlmLayout& L, int nj, const double* pp, const double* t, int k,
                                         double* w, double* tg, double* w2, double* tg2) {
  // stage lookup: stages have at most two sizes (torque stages first)
  int i = 0, off = 0;
  const int s0 = L.x_off[1] - L.x_off[0];
  if (L.tau_nodes > 0 && L.tau_nodes < L.nodes) {
    const int sw = L.x_off[L.tau_nodes];
    if (k < sw) { i = k / s0; off = k - i * s0; }
    else {
      const int s1 = L.x_off[L.tau_nodes + 1] - L.x_off[L.tau_nodes];
      i = L.tau_nodes + (k - sw) / s1; off = (k - sw) - (i - L.tau_nodes) * s1;
    }
  } else { i = k / s0; off = k - i * s0; }
  *w2 = 0.0; *tg2 = 0.0;
  if (off < L.ndx) { *w = pp[L.p_Q + off]; *tg = t[off]; }
  else {
    const int j = off - L.ndx;
    *w = pp[L.p_R + j]; *tg = t[L.ndx + j];
    if (L.dynamics == PLM_WHOLE_BODY_RNEA && i == 0 && L.tau_nodes > 0 && j >= L.tau_idx) {   // (tau_0 - tau_prev)^T W (.)
      *w2 = pp[L.p_W + j - L.tau_idx]; *tg2 = pp[L.p_tau_prev + j - L.tau_idx];
    }
  }
  (void)nj;
}

// f_data(x,p) -> f, grad_f.  One CTA per (instance, trial); x_eff = x + alpha * dx when dx != nullptr.
__global__ void objective_kernel(DeviceTables tab, const double* __restrict__ x, const double* __restrict__ dxs,
                                 const double* __restrict__ alphas, int t0, int ntrial, int ntot, const int* __restrict__ accepted,
                                 const double* __restrict__ p,
                                 const double* __restrict__ tgt, int tgt_ld, int batch,
                                 double* __restrict__ f, double* __restrict__ grad, double* __restrict__ gdot) {
  const PlmLayout& L = *tab.layout;
  const int b = blockIdx.x / ntrial, tr = t0 + blockIdx.x % ntrial;
  if (accepted && accepted[b]) return;
  const double* pp = p + (size_t)b * L.np;
  const double* t = tgt + (size_t)b * tgt_ld;
  const double* xb = x + (size_t)b * L.n;
  const double alpha = alphas ? alphas[tr] : 0.0;
  double acc = 0.0, dacc = 0.0;
  for (int k = threadIdx.x; k < L.n; k += blockDim.x) {
    double w, tg, w2, tg2;
    var_weight_target(L, tab.model->nj, pp, t, k, &w, &tg, &w2, &tg2);
    double xv = xb[k];
    if (dxs) xv += alpha * dxs[(size_t)b * L.n + k];
    double e = xv - tg, e2 = xv - tg2;
    acc += w * e * e + w2 * e2 * e2;
    const double gk = 2.0 * w * e + 2.0 * w2 * e2;
    if (grad) grad[(size_t)b * L.n + k] = gk;
    if (gdot) dacc += gk * dxs[(size_t)b * L.n + k];
  }
  __shared__ double red[32], red2[32];
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_down_sync(0xffffffffu, acc, o);
    dacc += __shfl_down_sync(0xffffffffu, dacc, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; red2[threadIdx.x >> 5] = dacc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0, v2 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { v += red[w]; v2 += red2[w]; }
    if (f) f[(size_t)b * ntot + tr] = v;
    if (gdot) gdot[b] = v2;
  }
}

// Constraint-violation metric of optimization/ocp.py:482-496 from stored rows: out[b] = (sum of squares, max)
__global__ void violation_kernel(DeviceTables tab, const double* __restrict__ g, const double* __restrict__ lbg,
                                 const double* __restrict__ ubg, int batch, double* __restrict__ out) {
  const PlmLayout& L = *tab.layout;
  const int b = blockIdx.x;
  double ss = 0.0, mx = 0.0;
  for (int r = threadIdx.x; r < L.m; r += blockDim.x) {
    const double gv = g[(size_t)b * L.m + r];
    const double v1 = fmax(0.0, lbg[(size_t)b * L.m + r] - gv), v2 = fmax(0.0, gv - ubg[(size_t)b * L.m + r]);
    ss += v1 * v1 + v2 * v2;
    mx = fmax(mx, fmax(v1, v2));
  }
  __shared__ double r1[32], r2[32];
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_down_sync(0xffffffffu, ss, o);
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = ss; r2[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, m2 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += r1[w]; m2 = fmax(m2, r2[w]); }
    out[2 * b] = s;
    out[2 * b + 1] = m2;
  }
}

// Armijo / filter acceptance of optimization/ocp.py:430-480, sequential over the trials [t0, t1) of one instance.
// state[b] = {f, g_metric, accepted, alpha_accepted, trials, armijo_metric, f0 (unused), viol_max}
__global__ void armijo_scan_kernel(DeviceTables tab, int batch, int t0, int t1, int ntot, const double* __restrict__ alphas,
                                   const double* __restrict__ ftr, const double* __restrict__ part,
                                   double* __restrict__ state, int* __restrict__ accepted) {
  const PlmLayout& L = *tab.layout;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double* st = state + 8 * (size_t)b;
  if (st[2] != 0.0) return;
  const double armijo_factor = 1e-4, g_max = 1e-3, g_min = 1e-5, gamma = 1e-5;
  double f = st[0], g_metric = st[1];
  const double armijo_metric = st[5];
  for (int t = t0; t < t1; ++t) {
    const double new_f = ftr[(size_t)b * ntot + t];
    double ss = 0.0, mx = 0.0;
    const double* pp = part + 2 * ((size_t)b * ntot + t) * L.nodes;
    for (int i = 0; i < L.nodes; ++i) { ss += pp[2 * i]; mx = fmax(mx, pp[2 * i + 1]); }
    const double new_g = sqrt(ss);
    bool acc = false;
    if (new_g > g_max) {
      if (new_g < (1.0 - gamma) * g_metric) acc = true;
    } else if (fmax(new_g, g_metric) < g_min && armijo_metric < 0.0) {
      if (new_f <= f + armijo_factor * armijo_metric) acc = true;
    } else if (new_f <= f - gamma * new_g || new_g < (1.0 - gamma) * g_metric) {
      acc = true;
    }
    f = new_f;
    g_metric = new_g;
    st[4] = (double)(t + 1);
    if (acc) {
      st[2] = 1.0;
      st[3] = alphas[t];
      st[7] = mx;
      accepted[b] = 1;
      break;
    }
  }
  st[0] = f;
  st[1] = g_metric;
}

// l = lbg - g, u = ubg - g  (optimization/ocp.py:393-394)
__global__ void bounds_shift_kernel(long long total, const double* __restrict__ g, const double* __restrict__ lbg,
                                    const double* __restrict__ ubg, double* __restrict__ l, double* __restrict__ u) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  l[i] = lbg[i] - g[i];
  u[i] = ubg[i] - g[i];
}

// state[b] = {f0, ||viol(x)||_2, 0, 0, 0, grad_f.dx, f0, max viol(x)}
__global__ void armijo_init_kernel(int batch, const double* __restrict__ f0, const double* __restrict__ viol,
                                   const double* __restrict__ gdot, double* __restrict__ state, int* __restrict__ accepted) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double* st = state + 8 * (size_t)b;
  st[0] = f0[b]; st[1] = sqrt(viol[2 * b]); st[2] = 0.0; st[3] = 0.0; st[4] = 0.0; st[5] = gdot[b]; st[6] = f0[b]; st[7] = viol[2 * b + 1];
  accepted[b] = 0;
}

// stats[b] = {qp iterations, qp status, accepted, step size, trials, f, g_metric, max violation}
__global__ void sqp_stats_kernel(int batch, const int* __restrict__ iters, const int* __restrict__ status,
                                 const double* __restrict__ state, double* __restrict__ stats) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* st = state + 8 * (size_t)b;
  double* o = stats + 8 * (size_t)b;
  o[0] = iters ? (double)iters[b] : 0.0; o[1] = status ? (double)status[b] : 0.0;
  o[2] = st[2]; o[3] = st[3]; o[4] = st[4]; o[5] = st[0]; o[6] = st[1]; o[7] = st[7];
}

// x_new = x + alpha_accepted * dx (or x when the line search failed); stats row of plm_sqp_step / info of plm_line_search
__global__ void armijo_apply_kernel(DeviceTables tab, int batch, const double* __restrict__ x, const double* __restrict__ dx,
                                    const double* __restrict__ state, double* __restrict__ x_new) {
  const PlmLayout& L = *tab.layout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)batch * L.n) return;
  const int b = (int)(idx / L.n);
  const double* st = state + 8 * (size_t)b;
  x_new[idx] = (st[2] != 0.0) ? x[idx] + st[3] * dx[idx] : x[idx];
}

__global__ void hess_diag_kernel(DeviceTables tab, const double* __restrict__ p, int batch, double* __restrict__ hess) {
  const PlmLayout& L = *tab.layout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)batch * L.n) return;
  const int b = (int)(idx / L.n), k = (int)(idx % L.n);
  const double* pp = p + (size_t)b * L.np;
  double w, tg, w2, tg2;
  var_weight_target(L, tab.model->nj, pp, pp /*unused targets*/, k, &w, &tg, &w2, &tg2);
  hess[idx] = 2.0 * w + 2.0 * w2;
}

}  // namespace plm
