// CUDA kernels of the evaluation side of the SQP iteration: node rows + Jacobian blocks, bounds, objective.
#pragma once
#include <cuda_runtime.h>

#include "plm_node_driver.cuh"

namespace plm {


// One warp per (instance, node).  Shared memory: [PlmModel | PlmLayout | per-warp NodeWs + row/J staging].
template <int KIND>
__global__ void __launch_bounds__(PLM_NODE_WARPS * 32)
node_eval_kernel(DeviceTables tab, const double* __restrict__ x, const double* __restrict__ p, int batch,
                 double* __restrict__ g, double* __restrict__ Jv, int want_jac, int ws_doubles) {
  extern __shared__ double smem[];
  PlmModel* sM = reinterpret_cast<PlmModel*>(smem);
  PlmLayout* sL = reinterpret_cast<PlmLayout*>(smem + (sizeof(PlmModel) + 7) / 8);
  double* wsbase = smem + (sizeof(PlmModel) + 7) / 8 + (sizeof(PlmLayout) + 7) / 8;
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
  if (item >= (long long)batch * L.nodes) return;
  const int b = (int)(item / L.nodes), node = (int)(item % L.nodes);
  NodeWs& ws = *reinterpret_cast<NodeWs*>(wsbase + (size_t)warp * ws_doubles);
  if (lane == 0) node_ws_bind(ws, L, wsbase + (size_t)warp * ws_doubles + (sizeof(NodeWs) + 7) / 8);
  __syncwarp();
  NodeArgs A;
  A.M = sM;
  A.L = sL;
  A.T = &L.types[L.node_type[node]];
  A.lut = tab.lut + A.T->lut_off;
  A.consts = tab.consts + A.T->const_off;
  A.xs = x + (size_t)b * L.n + L.x_off[node];
  A.p = p + (size_t)b * L.np;
  A.node = node;
  A.dt = node_dt(L, A.p, node);
  A.want_jac = want_jac;
  WarpExec ex;
  ex.lane = lane;
  node_eval_body<KIND>(ex, ws, A);
  // coalesced write-out
  const PlmNodeType& T = *A.T;
  double* go = g + (size_t)b * L.m + L.row_off[node];
  for (int r = lane; r < T.nrows; r += 32) go[r] = ws.g[r];
  if (node == 0) {   // DX_0 == 0 rows (optimization/ocp.py:109)
    double* g0 = g + (size_t)b * L.m;
    for (int r = lane; r < L.ndx; r += 32) g0[r] = A.xs[r];
  }
  if (want_jac) {
    double* Jo = Jv + (size_t)b * L.nnz + L.nnz_off[node];
    for (int e = lane; e < T.nnz; e += 32) Jo[e] = ws.J[e];
    if (node == 0) {
      double* J0 = Jv + (size_t)b * L.nnz;
      for (int e = lane; e < L.ndx; e += 32) J0[e] = 1.0;
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

// SE(3) logarithm of M0^-1 M1 from positions and unit quaternions [x,y,z,w]  (pin.difference, base part)
__device__ inline void se3_difference(const double* p0, const double* q0, const double* p1, const double* q1, double* nu) {
  // relative quaternion qr = conj(q0) * q1
  double ax = -q0[0], ay = -q0[1], az = -q0[2], aw = q0[3];
  double bx = q1[0], by = q1[1], bz = q1[2], bw = q1[3];
  double qr[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                  aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz};
  if (qr[3] < 0) { qr[0] = -qr[0]; qr[1] = -qr[1]; qr[2] = -qr[2]; qr[3] = -qr[3]; }
  double n = sqrt(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2]);
  double w[3];
  if (n < 1e-12) {
    double s = 2.0 / qr[3];
    w[0] = s * qr[0]; w[1] = s * qr[1]; w[2] = s * qr[2];
  } else {
    double s = 2.0 * atan2(n, qr[3]) / n;
    w[0] = s * qr[0]; w[1] = s * qr[1]; w[2] = s * qr[2];
  }
  double R0[9], d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, pl[3];
  quat_to_R(q0, R0);
  matTvec3(R0, d, pl);
  double t2 = dot3(w, w), beta;
  if (t2 < 1e-2) beta = 1.0 / 12 + t2 / 720 + t2 * t2 / 30240 + t2 * t2 * t2 / 1209600;
  else {
    double t = sqrt(t2);
    beta = 1.0 / t2 - (1.0 + cos(t)) / (2.0 * t * sin(t));
  }
  // Vinv p = p - 0.5 w x p + beta w x (w x p)
  double wp[3], wwp[3];
  cross3(w, pl, wp);
  cross3(w, wp, wwp);
  for (int i = 0; i < 3; ++i) nu[i] = pl[i] - 0.5 * wp[i] + beta * wwp[i];
  nu[3] = w[0]; nu[4] = w[1]; nu[5] = w[2];
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
                                 const double* __restrict__ alphas, int ntrial, const double* __restrict__ p,
                                 const double* __restrict__ tgt, int tgt_ld, int batch,
                                 double* __restrict__ f, double* __restrict__ grad) {
  const PlmLayout& L = *tab.layout;
  const int b = blockIdx.x / ntrial, tr = blockIdx.x % ntrial;
  const double* pp = p + (size_t)b * L.np;
  const double* t = tgt + (size_t)b * tgt_ld;
  const double* xb = x + (size_t)b * L.n;
  const double alpha = alphas ? alphas[tr] : 0.0;
  double acc = 0.0;
  for (int k = threadIdx.x; k < L.n; k += blockDim.x) {
    double w, tg, w2, tg2;
    var_weight_target(L, tab.model->nj, pp, t, k, &w, &tg, &w2, &tg2);
    double xv = xb[k];
    if (dxs) xv += alpha * dxs[(size_t)b * L.n + k];
    double e = xv - tg, e2 = xv - tg2;
    acc += w * e * e + w2 * e2 * e2;
    if (grad) grad[(size_t)b * L.n + k] = 2.0 * w * e + 2.0 * w2 * e2;
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && f) f[(size_t)b * ntrial + tr] = v;
  }
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
