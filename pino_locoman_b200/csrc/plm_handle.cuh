// Library-internal handle: host tables, device copies, workspaces.  One handle per (GPU, stream).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/pino_locoman_b200.h"
#include "plm_host.h"
#include "plm_types.h"
#include "plm_qp_types.h"

#define PLM_LS_TRIALS 14   // a = 1, 1/2, ... while a > 1e-4 (optimization/ocp.py:433-435,448)

namespace plm {
// Line-search and SQP-step workspaces (device), sized for max_batch.
struct LsWork {
  double* alphas = nullptr;   // [PLM_LS_TRIALS]
  double* ftr = nullptr;      // [B][PLM_LS_TRIALS] objective at the trials
  double* part = nullptr;     // [B][PLM_LS_TRIALS][nodes][2] violation partials
  double* state = nullptr;    // [B][8]
  double* viol = nullptr;     // [B][2]
  double* f0 = nullptr;       // [B]
  double* gdot = nullptr;     // [B]
  int* accepted = nullptr;    // [B]
  // plm_sqp_step buffers
  double *grad = nullptr, *J = nullptr, *g = nullptr, *lbg = nullptr, *ubg = nullptr, *l = nullptr, *u = nullptr, *hess = nullptr, *dx = nullptr;
  int *iters = nullptr, *status = nullptr;
};
}  // namespace plm

struct plm_probe;

struct plm_handle {
  plm::HostTables host;
  plm_ocp_desc ocp;
  int max_batch = 0;
  std::string error;
  PlmModel* d_model = nullptr;
  PlmLayout* d_layout = nullptr;
  int16_t* d_lut = nullptr;
  PlmConstEntry* d_consts = nullptr;
  plm::DeviceTables tab;
  double* d_tgt = nullptr;   // [max_batch][tgt_ld] dx_des | u_des
  int tgt_ld = 0;
  // per-warp workspace, warps (= node evaluations) per CTA (chosen to maximise the resident warps per SM) and dynamic
  // shared memory of the node kernel: [0] evaluation launches, [1] line-search trial launches (these stage x + alpha dx)
  int node_ws_doubles[2] = {0, 0};
  int node_warps[2] = {PLM_NODE_WARPS, PLM_NODE_WARPS};
  size_t node_smem[2] = {0, 0};
  long long launches = 0;
  // QP workspaces (plm_qp.cu)
  plm::QpWork qp;
  int qp_factor_doubles = 0;
  int* d_qp_fail = nullptr;   // [max_batch] stage index (+1) of a non-positive Cholesky pivot, 0 = ok
  size_t smem_scale = 0, smem_factor = 0, smem_admm = 0, smem_admm_lat = 0;
  // line search / SQP step workspaces, timing
  plm::LsWork ls;
  // lazily created two-node probe problems backing the Dynamics* entry points (one per formulation)
  plm_probe* probes[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int is_probe = 0;
  int qp_setup_batch = 0;   // instances [0, qp_setup_batch) have had their osqp setup (zero iterates, setup-time row scaling)
  int ls_alloc_done = 0;
  int sqp_alloc_done = 0;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double* d_mpc_dts = nullptr;   // [PLM_MAXNODES] horizon step sizes of plm_mpc_step
  double* d_frame_plc = nullptr; // [24] frame / base-frame placements of plm_frame_kinematics
  int num_sms = 148;             // multiprocessors of the device (kernel variant selection)
};

int plm_setup_node_kernels(plm_handle* h);
int plm_launch_node_eval(plm_handle* h, const double* x, const double* p, int batch, double* g, double* J, int want_jac, cudaStream_t s);
int plm_launch_bounds(plm_handle* h, const double* p, int batch, double* lbg, double* ubg, cudaStream_t s);
int plm_launch_targets(plm_handle* h, const double* p, int batch, cudaStream_t s);
int plm_launch_objective(plm_handle* h, const double* x, const double* dx, const double* alphas, int ntrial, const double* p,
                         int batch, double* f, double* grad, cudaStream_t s);
int plm_launch_hess_diag(plm_handle* h, const double* p, int batch, double* hess, cudaStream_t s);
int plm_launch_node_trials(plm_handle* h, const double* x, const double* p, int batch, double* g, double* J, int want_jac,
                           const void* trial_args, cudaStream_t s);
int plm_line_search_impl(plm_handle* h, const double* x, const double* p, const double* dx, int batch, const double* g,
                         const double* lbg, const double* ubg, double* x_new, cudaStream_t s);
int plm_launch_bounds_shift(plm_handle* h, int batch, const double* g, const double* lbg, const double* ubg, double* l, double* u, cudaStream_t s);
int plm_launch_stats(plm_handle* h, int batch, const int* iters, const int* status, double* stats, cudaStream_t s);
void plm_dyn_free(plm_handle* h);
int plm_qp_alloc(plm_handle* h);
int plm_qp_setup_impl(plm_handle* h, int first, int count, const double* d_hess, cudaStream_t s);
int plm_qp_update_impl(plm_handle* h, int batch, const double* d_hess, const double* d_q, const double* d_J, const double* d_l, const double* d_u, cudaStream_t s);
int plm_qp_solve_impl(plm_handle* h, int batch, double* d_dx, int* d_iters, int* d_status, cudaStream_t s);
void plm_qp_free(plm_handle* h);

#define PLM_LAUNCH_CHECK(h)                                                    \
  do {                                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    (h)->launches++;                                                           \
    if (_e != cudaSuccess) {                                                   \
      (h)->error = std::string("kernel launch: ") + cudaGetErrorString(_e);    \
      return 5;                                                                \
    }                                                                          \
  } while (0)
