// C ABI of libpinolocoman_b200.so (see include/pino_locoman_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pino_locoman_b200.h"
#include "plm_handle.cuh"
#include "plm_node_driver.cuh"

using namespace plm;

#define PLM_CHECK_CUDA(h, expr)                                                           \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                    \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

void plm_sqp_free(plm_handle* h);

// DFMA throughput microbenchmark: eight independent chains per thread (plm_fp64_peak)
__global__ void fp64_peak_kernel(int iters, double m, double* sink) {
  double a0 = threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) *sink = r;      // never true: keeps the chains alive
}

extern "C" {

void plm_fill_default_ocp_desc(plm_ocp_desc* d, int32_t dynamics, int32_t nodes) {
  memset(d, 0, sizeof(*d));
  d->dynamics = dynamics;
  d->nodes = nodes;
  d->tau_nodes = 3;            // ocp_args.py:16
  d->mu = 0.7;                 // optimization/ocp.py:103
  d->osqp_max_iter = 100;      // optimization/ocp.py:267-273
  d->osqp_alpha = 1.4;
  d->osqp_rho = 2e-2;
  d->osqp_check_termination = 25;   // osqp defaults below
  d->osqp_scaling = 10;
  d->osqp_sigma = 1e-6;
  d->osqp_eps_abs = 1e-3;
  d->osqp_eps_rel = 1e-3;
  d->osqp_eps_prim_inf = 1e-4;
  d->osqp_eps_dual_inf = 1e-4;
  d->include_base = 1;         // ocp_args.py:3-11
  d->include_acc = 1;          // ocp_args.py:17
}

void plm_abi_struct_sizes(int32_t out[3]) {
  out[0] = (int32_t)sizeof(plm_robot_desc);
  out[1] = (int32_t)sizeof(plm_ocp_desc);
  out[2] = (int32_t)sizeof(plm_dims);
}

int plm_create(const plm_robot_desc* robot, const plm_ocp_desc* ocp, int32_t max_batch, plm_handle** out) {
  *out = nullptr;
  plm_handle* h = new plm_handle();
  h->ocp = *ocp;
  h->max_batch = max_batch;
  *out = h;
  if (!build_tables(*robot, *ocp, h->host)) { h->error = h->host.error; return 2; }
  if (max_batch == 0) return 0;   // layout-only handle: tables and layout queries, no device work
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    h->error = "no CUDA device: pino_locoman_b200 has no CPU path";
    return 3;
  }
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      h->num_sms = sms;
  }
  PLM_CHECK_CUDA(h, cudaMalloc(&h->d_model, sizeof(PlmModel)));
  PLM_CHECK_CUDA(h, cudaMalloc(&h->d_layout, sizeof(PlmLayout)));
  PLM_CHECK_CUDA(h, cudaMalloc(&h->d_lut, h->host.lut.size() * sizeof(int16_t)));
  PLM_CHECK_CUDA(h, cudaMalloc(&h->d_consts, std::max<size_t>(1, h->host.consts.size()) * sizeof(PlmConstEntry)));
  PLM_CHECK_CUDA(h, cudaMemcpy(h->d_model, &h->host.model, sizeof(PlmModel), cudaMemcpyHostToDevice));
  PLM_CHECK_CUDA(h, cudaMemcpy(h->d_layout, &h->host.layout, sizeof(PlmLayout), cudaMemcpyHostToDevice));
  PLM_CHECK_CUDA(h, cudaMemcpy(h->d_lut, h->host.lut.data(), h->host.lut.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
  if (!h->host.consts.empty())
    PLM_CHECK_CUDA(h, cudaMemcpy(h->d_consts, h->host.consts.data(), h->host.consts.size() * sizeof(PlmConstEntry), cudaMemcpyHostToDevice));
  h->tab.model = h->d_model;
  h->tab.layout = h->d_layout;
  h->tab.lut = h->d_lut;
  h->tab.consts = h->d_consts;
  const PlmLayout& L = h->host.layout;
  h->tgt_ld = L.ndx + L.types[L.node_type[0]].nu;
  PLM_CHECK_CUDA(h, cudaMalloc(&h->d_tgt, (size_t)max_batch * h->tgt_ld * sizeof(double)));
  for (int mode = 0; mode < 2; ++mode) {
    h->node_ws_doubles[mode] = (int)node_ws_doubles(L, h->host.model.nv, L.nf, h->host.model.nbody, false, mode == 1);
    const size_t tables = ((sizeof(PlmModel) + 7) / 8 + (sizeof(PlmLayout) + 7) / 8) * 8;
    const size_t per_warp = (size_t)h->node_ws_doubles[mode] * 8, cap = 227 * 1024;
    int best_w = 0, best = 0;
    for (int w = 1; w <= PLM_NODE_WARPS; ++w) {
      const size_t need = tables + w * per_warp + 1024;   // + per-CTA reservation
      if (need > cap) break;
      // CTAs per SM: by shared memory and by registers (128 per thread under __launch_bounds__(256, 2): 16 warps per SM)
      const int ctas = std::min((int)(cap / need), 16 / w);
      const int resident = ctas * w;
      if (resident >= best) { best = resident; best_w = w; }
    }
    if (best_w == 0) { h->error = "node workspace exceeds shared memory"; return 6; }
    h->node_warps[mode] = best_w;
    h->node_smem[mode] = tables + best_w * per_warp;
  }
  int rc = plm_setup_node_kernels(h);
  if (rc) return rc;
  rc = plm_qp_alloc(h);
  if (rc) return rc;
  return 0;
}

void plm_destroy(plm_handle* h) {
  if (!h) return;
  if (h->max_batch == 0) { delete h; return; }
  plm_dyn_free(h);
  plm_qp_free(h);
  plm_sqp_free(h);
  cudaFree(h->d_model);
  cudaFree(h->d_layout);
  cudaFree(h->d_lut);
  cudaFree(h->d_consts);
  cudaFree(h->d_tgt);
  cudaFree(h->d_mpc_dts);
  delete h;
}

const char* plm_last_error(const plm_handle* h) { return h ? h->error.c_str() : "null handle"; }

int plm_get_dims(const plm_handle* h, plm_dims* d) {
  const PlmModel& M = h->host.model;
  const PlmLayout& L = h->host.layout;
  d->nq = M.nq; d->nv = M.nv; d->nj = M.nj; d->nf = L.nf;
  d->nx = L.nx; d->ndx = L.ndx; d->n = L.n; d->m = L.m; d->np = L.np; d->nnz = L.nnz;
  d->nodes = L.nodes;
  d->kkt_factor_doubles = h->host.qp.fac_total;      // host table: also valid for layout-only handles
  return 0;
}

int plm_stage_offsets(const plm_handle* h, int32_t* x_off, int32_t* nu, int32_t* row_off) {
  const PlmLayout& L = h->host.layout;
  for (int i = 0; i <= L.nodes; ++i) x_off[i] = L.x_off[i];
  x_off[L.nodes + 1] = L.n;
  for (int i = 0; i < L.nodes; ++i) nu[i] = L.types[L.node_type[i]].nu;
  for (int i = 0; i <= L.nodes; ++i) row_off[i] = L.row_off[i];
  row_off[L.nodes + 1] = L.m;
  return 0;
}

int plm_param_offsets(const plm_handle* h, int32_t* off) {
  const PlmLayout& L = h->host.layout;
  const int32_t v[16] = {L.p_x_init, L.p_dt_min, L.p_dt_max, L.p_contact, L.p_swing, L.p_n_contacts, L.p_swing_period,
                         L.p_swing_height, L.p_swing_vel, L.p_Q, L.p_R, L.p_base_vel, L.p_ext_force, L.p_arm_vel,
                         L.p_tau_prev, L.p_W};
  for (int i = 0; i < 16; ++i) off[i] = v[i];
  return 0;
}

int plm_jac_pattern(const plm_handle* h, int32_t* rows, int32_t* cols) {
  memcpy(rows, h->host.pat_rows.data(), h->host.pat_rows.size() * sizeof(int32_t));
  memcpy(cols, h->host.pat_cols.data(), h->host.pat_cols.size() * sizeof(int32_t));
  return 0;
}

static int check_batch(plm_handle* h, int batch) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  return 0;
}

int plm_g_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch, double* d_g, double* d_lbg,
               double* d_ubg, void* stream) {
  if (int rc = check_batch(h, batch)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = plm_launch_node_eval(h, d_x, d_p, batch, d_g, nullptr, 0, s)) return rc;
  if (d_lbg && d_ubg)
    if (int rc = plm_launch_bounds(h, d_p, batch, d_lbg, d_ubg, s)) return rc;
  return 0;
}

int plm_f_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch, double* d_f, double* d_grad_f, void* stream) {
  if (int rc = check_batch(h, batch)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = plm_launch_targets(h, d_p, batch, s)) return rc;
  return plm_launch_objective(h, d_x, nullptr, nullptr, 1, d_p, batch, d_f, d_grad_f, s);
}

int plm_sqp_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch, double* d_grad_f, double* d_J,
                 double* d_g, double* d_lbg, double* d_ubg, void* stream) {
  if (int rc = check_batch(h, batch)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = plm_launch_node_eval(h, d_x, d_p, batch, d_g, d_J, 1, s)) return rc;
  if (d_lbg && d_ubg)
    if (int rc = plm_launch_bounds(h, d_p, batch, d_lbg, d_ubg, s)) return rc;
  if (d_grad_f) {
    if (int rc = plm_launch_targets(h, d_p, batch, s)) return rc;
    if (int rc = plm_launch_objective(h, d_x, nullptr, nullptr, 1, d_p, batch, nullptr, d_grad_f, s)) return rc;
  }
  return 0;
}

int plm_hess_diag(plm_handle* h, const double* d_p, int32_t batch, double* d_hess, void* stream) {
  if (int rc = check_batch(h, batch)) return rc;
  return plm_launch_hess_diag(h, d_p, batch, d_hess, (cudaStream_t)stream);
}

int64_t plm_launch_count(const plm_handle* h) { return h->launches; }

int plm_fp64_peak(double* tflops) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 3;
  double* sink = nullptr;
  if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) return 7;
  const int blocks = sms * 4, threads = 512, iters = 1 << 15;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {       // first repetition warms up
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(iters, 1.0000001, sink);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return 5; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads / ((double)best * 1e-3) * 1e-12;
  return 0;
}

}  // extern "C"
