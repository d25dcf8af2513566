// Kernel launchers for the evaluation side (node rows / Jacobian, bounds, objective).
#include "plm_handle.cuh"
#include "plm_kernels.cuh"
#include <string.h>

using namespace plm;

template <int KIND, bool NOBASE>
static int setup_one(plm_handle* h) {
  cudaError_t e = cudaFuncSetAttribute(node_eval_kernel<KIND, NOBASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per function, shared by all handles
  if (e != cudaSuccess) { h->error = std::string("node kernel shared memory: ") + cudaGetErrorString(e); return 6; }
  return 0;
}

int plm_setup_node_kernels(plm_handle* h) {
  if (h->node_smem[0] > 227 * 1024 || h->node_smem[1] > 227 * 1024) { h->error = "node workspace exceeds shared memory"; return 6; }
  const bool nb = h->host.layout.nobase != 0;
  switch (h->host.layout.dynamics) {
    case PLM_CENTROIDAL_VEL: return nb ? setup_one<PLM_CENTROIDAL_VEL, true>(h) : setup_one<PLM_CENTROIDAL_VEL, false>(h);
    case PLM_CENTROIDAL_ACC: return nb ? setup_one<PLM_CENTROIDAL_ACC, true>(h) : setup_one<PLM_CENTROIDAL_ACC, false>(h);
    case PLM_WHOLE_BODY_ACC: return nb ? setup_one<PLM_WHOLE_BODY_ACC, true>(h) : setup_one<PLM_WHOLE_BODY_ACC, false>(h);
    case PLM_WHOLE_BODY_ABA: return setup_one<PLM_WHOLE_BODY_ABA, false>(h);
    default: return setup_one<PLM_WHOLE_BODY_RNEA, false>(h);
  }
}

int plm_launch_node_eval(plm_handle* h, const double* x, const double* p, int batch, double* g, double* J, int want_jac, cudaStream_t s) {
  TrialArgs tr;
  memset(&tr, 0, sizeof(tr));
  return plm_launch_node_trials(h, x, p, batch, g, J, want_jac, &tr, s);
}

int plm_launch_node_trials(plm_handle* h, const double* x, const double* p, int batch, double* g, double* J, int want_jac,
                           const void* trial_args, cudaStream_t s) {
  const TrialArgs tr = *reinterpret_cast<const TrialArgs*>(trial_args);
  const PlmLayout& L = h->host.layout;
  const long long items = (long long)batch * L.nodes * (tr.part ? tr.ntrial : 1);
  const int mode = tr.part ? 1 : 0;      // line-search trial launches carry the staged trial point
  const int blocks = (int)((items + h->node_warps[mode] - 1) / h->node_warps[mode]);
  const dim3 grid(blocks), block(h->node_warps[mode] * 32);
#define PLM_NODE_LAUNCH(KIND, NB) node_eval_kernel<KIND, NB><<<grid, block, h->node_smem[mode], s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles[mode], tr)
  const bool nb = L.nobase != 0;
  switch (L.dynamics) {
    case PLM_CENTROIDAL_VEL: if (nb) PLM_NODE_LAUNCH(PLM_CENTROIDAL_VEL, true); else PLM_NODE_LAUNCH(PLM_CENTROIDAL_VEL, false); break;
    case PLM_CENTROIDAL_ACC: if (nb) PLM_NODE_LAUNCH(PLM_CENTROIDAL_ACC, true); else PLM_NODE_LAUNCH(PLM_CENTROIDAL_ACC, false); break;
    case PLM_WHOLE_BODY_ACC: if (nb) PLM_NODE_LAUNCH(PLM_WHOLE_BODY_ACC, true); else PLM_NODE_LAUNCH(PLM_WHOLE_BODY_ACC, false); break;
    case PLM_WHOLE_BODY_ABA: PLM_NODE_LAUNCH(PLM_WHOLE_BODY_ABA, false); break;
    default: PLM_NODE_LAUNCH(PLM_WHOLE_BODY_RNEA, false); break;
  }
#undef PLM_NODE_LAUNCH
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_bounds(plm_handle* h, const double* p, int batch, double* lbg, double* ubg, cudaStream_t s) {
  const PlmLayout& L = h->host.layout;
  const long long threads = (long long)batch * L.nodes * 32;
  bounds_kernel<<<(int)((threads + 127) / 128), 128, 0, s>>>(h->tab, p, batch, lbg, ubg);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_targets(plm_handle* h, const double* p, int batch, cudaStream_t s) {
  targets_kernel<<<(batch + 63) / 64, 64, 0, s>>>(h->tab, p, batch, h->d_tgt, h->tgt_ld);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_objective(plm_handle* h, const double* x, const double* dx, const double* alphas, int ntrial, const double* p,
                         int batch, double* f, double* grad, cudaStream_t s) {
  objective_kernel<<<batch * ntrial, 256, 0, s>>>(h->tab, x, dx, alphas, 0, ntrial, ntrial, nullptr, p, h->d_tgt, h->tgt_ld, batch, f, grad, nullptr);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

// ---- Armijo line search (optimization/ocp.py:430-480) on precomputed g(x), lbg, ubg
int plm_line_search_impl(plm_handle* h, const double* x, const double* p, const double* dx, int batch, const double* g,
                         const double* lbg, const double* ubg, double* x_new, cudaStream_t s) {
  const PlmLayout& L = h->host.layout;
  plm::LsWork& W = h->ls;
  const int T = PLM_LS_TRIALS;
  violation_kernel<<<batch, 256, 0, s>>>(h->tab, g, lbg, ubg, batch, W.viol);
  PLM_LAUNCH_CHECK(h);
  // f(x) and the Armijo metric grad_f . dx
  objective_kernel<<<batch, 256, 0, s>>>(h->tab, x, dx, nullptr, 0, 1, 1, nullptr, p, h->d_tgt, h->tgt_ld, batch, W.f0, nullptr, W.gdot);
  PLM_LAUNCH_CHECK(h);
  armijo_init_kernel<<<(batch + 127) / 128, 128, 0, s>>>(batch, W.f0, W.viol, W.gdot, W.state, W.accepted);
  PLM_LAUNCH_CHECK(h);
  // trials are evaluated in rounds: every instance tries the full step, then the instances not yet accepted try the
  // next candidates (most accept within the first two); the remaining ones evaluate all remaining step sizes at once
  const int ranges[5] = {0, 1, 2, 4 < T ? 4 : T, T};
  for (int k = 0; k < 4; ++k) {
    if (ranges[k + 1] <= ranges[k]) continue;
    const int t0 = ranges[k], cnt = ranges[k + 1] - ranges[k];
    TrialArgs tr;
    tr.dxs = dx; tr.alphas = W.alphas; tr.t0 = t0; tr.ntrial = cnt; tr.ntot = T;
    tr.accepted = (k == 0) ? nullptr : W.accepted; tr.lbg = lbg; tr.ubg = ubg; tr.part = W.part;
    if (int rc = plm_launch_node_trials(h, x, p, batch, nullptr, nullptr, 0, &tr, s)) return rc;
    objective_kernel<<<batch * cnt, 256, 0, s>>>(h->tab, x, dx, W.alphas, t0, cnt, T, tr.accepted, p, h->d_tgt, h->tgt_ld, batch, W.ftr, nullptr, nullptr);
    PLM_LAUNCH_CHECK(h);
    armijo_scan_kernel<<<(batch + 127) / 128, 128, 0, s>>>(h->tab, batch, t0, t0 + cnt, T, W.alphas, W.ftr, W.part, W.state, W.accepted);
    PLM_LAUNCH_CHECK(h);
  }
  const long long tot = (long long)batch * L.n;
  armijo_apply_kernel<<<(int)((tot + 255) / 256), 256, 0, s>>>(h->tab, batch, x, dx, W.state, x_new);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_bounds_shift(plm_handle* h, int batch, const double* g, const double* lbg, const double* ubg, double* l, double* u, cudaStream_t s) {
  const long long tot = (long long)batch * h->host.layout.m;
  bounds_shift_kernel<<<(int)((tot + 255) / 256), 256, 0, s>>>(tot, g, lbg, ubg, l, u);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_stats(plm_handle* h, int batch, const int* iters, const int* status, double* stats, cudaStream_t s) {
  sqp_stats_kernel<<<(batch + 127) / 128, 128, 0, s>>>(batch, iters, status, h->ls.state, stats);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_hess_diag(plm_handle* h, const double* p, int batch, double* hess, cudaStream_t s) {
  const long long tot = (long long)batch * h->host.layout.n;
  hess_diag_kernel<<<(int)((tot + 255) / 256), 256, 0, s>>>(h->tab, p, batch, hess);
  PLM_LAUNCH_CHECK(h);
  return 0;
}
