// Kernel launchers for the evaluation side (node rows / Jacobian, bounds, objective).
#include "plm_handle.cuh"
#include "plm_kernels.cuh"

using namespace plm;

template <int KIND>
static int setup_one(plm_handle* h) {
  cudaError_t e = cudaFuncSetAttribute(node_eval_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->node_smem);
  if (e != cudaSuccess) { h->error = std::string("node kernel shared memory: ") + cudaGetErrorString(e); return 6; }
  return 0;
}

int plm_setup_node_kernels(plm_handle* h) {
  if (h->node_smem > 227 * 1024) { h->error = "node workspace exceeds shared memory"; return 6; }
  switch (h->host.layout.dynamics) {
    case PLM_CENTROIDAL_VEL: return setup_one<PLM_CENTROIDAL_VEL>(h);
    case PLM_CENTROIDAL_ACC: return setup_one<PLM_CENTROIDAL_ACC>(h);
    case PLM_WHOLE_BODY_ACC: return setup_one<PLM_WHOLE_BODY_ACC>(h);
    case PLM_WHOLE_BODY_ABA: return setup_one<PLM_WHOLE_BODY_ABA>(h);
    default: return setup_one<PLM_WHOLE_BODY_RNEA>(h);
  }
}

int plm_launch_node_eval(plm_handle* h, const double* x, const double* p, int batch, double* g, double* J, int want_jac, cudaStream_t s) {
  const PlmLayout& L = h->host.layout;
  const long long items = (long long)batch * L.nodes;
  const int blocks = (int)((items + h->node_warps - 1) / h->node_warps);
  const dim3 grid(blocks), block(h->node_warps * 32);
  switch (L.dynamics) {
    case PLM_CENTROIDAL_VEL: node_eval_kernel<PLM_CENTROIDAL_VEL><<<grid, block, h->node_smem, s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles); break;
    case PLM_CENTROIDAL_ACC: node_eval_kernel<PLM_CENTROIDAL_ACC><<<grid, block, h->node_smem, s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles); break;
    case PLM_WHOLE_BODY_ACC: node_eval_kernel<PLM_WHOLE_BODY_ACC><<<grid, block, h->node_smem, s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles); break;
    case PLM_WHOLE_BODY_ABA: node_eval_kernel<PLM_WHOLE_BODY_ABA><<<grid, block, h->node_smem, s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles); break;
    default: node_eval_kernel<PLM_WHOLE_BODY_RNEA><<<grid, block, h->node_smem, s>>>(h->tab, x, p, batch, g, J, want_jac, h->node_ws_doubles); break;
  }
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
  objective_kernel<<<batch * ntrial, 256, 0, s>>>(h->tab, x, dx, alphas, ntrial, p, h->d_tgt, h->tgt_ld, batch, f, grad);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_launch_hess_diag(plm_handle* h, const double* p, int batch, double* hess, cudaStream_t s) {
  const long long tot = (long long)batch * h->host.layout.n;
  hess_diag_kernel<<<(int)((tot + 255) / 256), 256, 0, s>>>(h->tab, p, batch, hess);
  PLM_LAUNCH_CHECK(h);
  return 0;
}
