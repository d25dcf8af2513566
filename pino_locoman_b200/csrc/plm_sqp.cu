// Line search and the fused SQP step: orchestration of the kernels on the caller's stream.
// Replaces the body of the loop at optimization/ocp.py:383-406 and _armijo_line_search (:430-480).
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: no-ops unless a profiler injects itself (nsys / ncu --nvtx)

#include "plm_handle.cuh"

using namespace plm;

// NVTX range per phase of the SQP iteration (tracing); closes on scope exit
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
struct NvtxPhases {   // consecutive sub-ranges; whatever is open is closed on every exit path
  bool open = false;
  void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
  ~NvtxPhases() { if (open) nvtxRangePop(); }
};

#define SQP_CUDA(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return 8;                                                                        \
    }                                                                                  \
  } while (0)

// allocate once; a failed call leaves the workspace marked not ready and the next call allocates what is missing
template <class T>
static cudaError_t need(T** p, size_t bytes) { return *p ? cudaSuccess : cudaMalloc(p, bytes); }

static int ensure_ls(plm_handle* h) {
  LsWork& W = h->ls;
  if (h->ls_alloc_done) return 0;
  const PlmLayout& L = h->host.layout;
  const size_t B = (size_t)h->max_batch;
  double al[PLM_LS_TRIALS];
  double a = 1.0;
  for (int t = 0; t < PLM_LS_TRIALS; ++t) { al[t] = a; a *= 0.5; }
  SQP_CUDA(h, need(&W.alphas, sizeof(al)));
  SQP_CUDA(h, cudaMemcpy(W.alphas, al, sizeof(al), cudaMemcpyHostToDevice));
  SQP_CUDA(h, need(&W.ftr, B * PLM_LS_TRIALS * 8));
  SQP_CUDA(h, need(&W.part, B * PLM_LS_TRIALS * L.nodes * 2 * 8));
  SQP_CUDA(h, need(&W.state, B * 8 * 8));
  SQP_CUDA(h, need(&W.viol, B * 2 * 8));
  SQP_CUDA(h, need(&W.f0, B * 8));
  SQP_CUDA(h, need(&W.gdot, B * 8));
  SQP_CUDA(h, need(&W.accepted, B * sizeof(int)));
  SQP_CUDA(h, need(&W.g, B * L.m * 8));
  SQP_CUDA(h, need(&W.lbg, B * L.m * 8));
  SQP_CUDA(h, need(&W.ubg, B * L.m * 8));
  h->ls_alloc_done = 1;
  return 0;
}

static int ensure_sqp(plm_handle* h) {
  if (int rc = ensure_ls(h)) return rc;
  LsWork& W = h->ls;
  if (h->sqp_alloc_done) return 0;
  const PlmLayout& L = h->host.layout;
  const size_t B = (size_t)h->max_batch;
  SQP_CUDA(h, need(&W.grad, B * L.n * 8));
  SQP_CUDA(h, need(&W.J, B * L.nnz * 8));
  SQP_CUDA(h, need(&W.l, B * L.m * 8));
  SQP_CUDA(h, need(&W.u, B * L.m * 8));
  SQP_CUDA(h, need(&W.hess, B * L.n * 8));
  SQP_CUDA(h, need(&W.dx, B * L.n * 8));
  SQP_CUDA(h, need(&W.iters, B * sizeof(int)));
  SQP_CUDA(h, need(&W.status, B * sizeof(int)));
  for (int i = 0; i < 5; ++i)
    if (!h->ev[i]) SQP_CUDA(h, cudaEventCreate(&h->ev[i]));
  h->sqp_alloc_done = 1;
  return 0;
}

void plm_sqp_free(plm_handle* h) {
  LsWork& W = h->ls;
  cudaFree(W.alphas); cudaFree(W.ftr); cudaFree(W.part); cudaFree(W.state); cudaFree(W.viol); cudaFree(W.f0); cudaFree(W.gdot);
  cudaFree(W.accepted); cudaFree(W.g); cudaFree(W.lbg); cudaFree(W.ubg); cudaFree(W.grad); cudaFree(W.J); cudaFree(W.l);
  cudaFree(W.u); cudaFree(W.hess); cudaFree(W.dx); cudaFree(W.iters); cudaFree(W.status);
  for (int i = 0; i < 5; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
}

extern "C" {

int plm_line_search(plm_handle* h, const double* d_x, const double* d_p, const double* d_dx, int32_t batch,
                    double* d_x_new, double* d_info, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  if (int rc = ensure_ls(h)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LsWork& W = h->ls;
  if (int rc = plm_launch_targets(h, d_p, batch, s)) return rc;
  if (int rc = plm_launch_node_eval(h, d_x, d_p, batch, W.g, nullptr, 0, s)) return rc;
  if (int rc = plm_launch_bounds(h, d_p, batch, W.lbg, W.ubg, s)) return rc;
  if (int rc = plm_line_search_impl(h, d_x, d_p, d_dx, batch, W.g, W.lbg, W.ubg, d_x_new, s)) return rc;
  if (d_info) {
    // info = {accepted, step, trials, g_metric}: gather from the state rows
    SQP_CUDA(h, cudaMemcpy2DAsync(d_info, 4 * 8, W.state + 2, 8 * 8, 3 * 8, batch, cudaMemcpyDeviceToDevice, s));
    SQP_CUDA(h, cudaMemcpy2DAsync(d_info + 3, 4 * 8, W.state + 1, 8 * 8, 8, batch, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

int plm_sqp_step(plm_handle* h, const double* d_x, const double* d_p, int32_t batch, double* d_x_new, double* d_stats, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  if (int rc = ensure_sqp(h)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  LsWork& W = h->ls;
  NvtxRange r_step("plm_sqp_step");
  NvtxPhases ph;
  SQP_CUDA(h, cudaEventRecord(h->ev[0], s));
  // ---- sqp_data(x, p)  (optimization/ocp.py:386)
  ph.next("sqp_data");
  if (int rc = plm_launch_targets(h, d_p, batch, s)) return rc;
  if (int rc = plm_launch_node_eval(h, d_x, d_p, batch, W.g, W.J, 1, s)) return rc;
  if (int rc = plm_launch_bounds(h, d_p, batch, W.lbg, W.ubg, s)) return rc;
  if (int rc = plm_launch_objective(h, d_x, nullptr, nullptr, 1, d_p, batch, nullptr, W.grad, s)) return rc;
  if (int rc = plm_launch_hess_diag(h, d_p, batch, W.hess, s)) return rc;
  SQP_CUDA(h, cudaEventRecord(h->ev[1], s));
  // ---- osqp update (optimization/ocp.py:391-395)
  ph.next("osqp_update");
  if (h->qp_setup_batch < batch) {
    // lazy osqp setup (optimization/ocp.py:305-313) of the instances that have not been set up yet: zero iterates and the
    // setup-time row scaling the first bound classification uses
    if (int rc = plm_qp_setup_impl(h, h->qp_setup_batch, batch - h->qp_setup_batch, W.hess, s)) return rc;
    h->qp_setup_batch = batch;
  }
  if (int rc = plm_launch_bounds_shift(h, batch, W.g, W.lbg, W.ubg, W.l, W.u, s)) return rc;
  if (int rc = plm_qp_update_impl(h, batch, W.hess, W.grad, W.J, W.l, W.u, s)) return rc;
  SQP_CUDA(h, cudaEventRecord(h->ev[2], s));
  // ---- osqp solve (optimization/ocp.py:401)
  ph.next("osqp_solve");
  if (int rc = plm_qp_solve_impl(h, batch, W.dx, W.iters, W.status, s)) return rc;
  SQP_CUDA(h, cudaEventRecord(h->ev[3], s));
  // ---- Armijo line search (optimization/ocp.py:406)
  ph.next("armijo_line_search");
  if (int rc = plm_line_search_impl(h, d_x, d_p, W.dx, batch, W.g, W.lbg, W.ubg, d_x_new, s)) return rc;
  if (d_stats)
    if (int rc = plm_launch_stats(h, batch, W.iters, W.status, d_stats, s)) return rc;
  SQP_CUDA(h, cudaEventRecord(h->ev[4], s));
  return 0;
}

int plm_last_phase_ms(plm_handle* h, double* ms4) {
  if (!h->sqp_alloc_done) { h->error = "plm_sqp_step has not run on this handle"; return 9; }
  SQP_CUDA(h, cudaEventSynchronize(h->ev[4]));
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    SQP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    ms4[i] = ms;
  }
  return 0;
}

}  // extern "C"
