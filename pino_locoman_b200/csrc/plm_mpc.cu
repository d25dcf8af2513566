// GPU-resident receding-horizon step: the body of the generic loop of run_mpc.py:127-143 for a
// batch of independent scenarios, without a host round trip between the parameter update, the SQP iteration and the
// state advance:
//   update_gait_sequence(t)      -> contact / swing schedules written into p        (utils/gait_sequence.py:5-77)
//   warm_start()                 -> forces of the previous solution reset to the contact-masked f_des, everything else
//                                   kept un-shifted                                  (ocp_whole_body_rnea.py:207-235)
//   solve()                      -> plm_sqp_step                                     (optimization/ocp.py:375-422)
//   x_init = integrate(x_init, DX_prev[1])                                           (run_mpc.py:141)
#include <cuda_runtime.h>

#include <string>

#include "../../include/pino_locoman_b200.h"
#include "plm_handle.cuh"
#include "plm_vec.cuh"

using namespace plm;

namespace {

// Python's float % for a positive period (utils/gait_sequence.py:41-42): C fmod keeps the sign of t, Python floors.
__device__ __forceinline__ double py_mod(double t, double period) {
  double r = fmod(t, period);
  if (r < 0.0) r += period;
  return r;
}

// One CTA per instance.  The schedule follows the reference operation order exactly (time accumulated node by node,
// floor-mod for the phases): the flags and phases are bit-identical to the host GaitSequence.
// x_mode: 0 = no warm start: x <- opti.initial() (DX = 0, U_i = u_des[:nu_i]: zero leading block and torques, unmasked
// f_des; optimization/ocp.py:159-163,193), what every solve() of the reference starts from when warm_start() is not
// called (run_mpc.py:131-132); 1 = warm_start(): forces <- contact-masked f_des, the rest of the previous solution
// kept; 2 = leave x alone (first step: the caller's initial point).
__global__ void mpc_prepare_kernel(const PlmLayout* __restrict__ Lp, int batch, int gait, double gait_period, double swing_period, int n_contacts,
                                   const double* __restrict__ dts, const double* __restrict__ t0, double t_add, double mass,
                                   int x_mode, double* __restrict__ x, double* __restrict__ p) {
  const PlmLayout& L = *Lp;
  const int b = blockIdx.x;
  if (b >= batch) return;
  double* pb = p + (size_t)b * L.np;
  double* xb = x + (size_t)b * L.n;
  const int N = L.nodes;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double t = (t0 ? t0[b] : 0.0) + t_add;
    for (int j = 0; j < i; ++j) t = t + dts[j];
    double contact[4] = {1.0, 1.0, 1.0, 1.0}, swing[4] = {0.0, 0.0, 0.0, 0.0};
    if (gait != 2) {
      const double gait_phase = py_mod(t, gait_period) / gait_period;
      const double swing_phase = py_mod(t, swing_period) / swing_period;
      int f0, f1 = -1;
      if (gait == 0) {                       // trot: FR + RL, then FL + RR
        if (gait_phase < 0.5) { f0 = 0; f1 = 3; } else { f0 = 1; f1 = 2; }
      } else {                               // walk: FL, RR, FR, RL
        f0 = gait_phase < 0.25 ? 1 : (gait_phase < 0.5 ? 2 : (gait_phase < 0.75 ? 0 : 3));
      }
      contact[f0] = 0.0; swing[f0] = swing_phase;
      if (f1 >= 0) { contact[f1] = 0.0; swing[f1] = swing_phase; }
    }
    for (int f = 0; f < 4; ++f) {
      pb[L.p_contact + 4 * i + f] = contact[f];
      pb[L.p_swing + 4 * i + f] = swing[f];
    }
    if (x_mode == 0) {
      double* xs = xb + L.x_off[i];
      const int len = L.x_off[i + 1] - L.x_off[i];
      for (int j = 0; j < len; ++j) xs[j] = 0.0;
      if (i == N - 1) for (int j = 0; j < L.ndx; ++j) xb[L.x_off[N] + j] = 0.0;
    }
    if (x_mode != 2) {
      // forces of node i <- f_des (z components 0.8 / 1.2 m g / n_contacts front / rear), masked by the contact flags
      // in a warm start; the external-force entries of f_des are zero
      const double fg = 9.81 * mass;
      const double fz_front = 0.8 * fg / (double)n_contacts, fz_rear = 1.2 * fg / (double)n_contacts;
      double* f = xb + L.x_off[i] + L.ndx + L.f_idx;
      for (int j = 0; j < L.nf; ++j) {
        double v = 0.0;
        if (j < 12 && (j % 3) == 2) v = ((j / 3) < 2 ? fz_front : fz_rear) * ((x_mode == 0 || contact[j / 3] != 0.0) ? 1.0 : 0.0);
        f[j] = v;
      }
    }
  }
  if (threadIdx.x == 0) {
    pb[L.p_n_contacts] = (double)n_contacts;
    pb[L.p_swing_period] = swing_period;
  }
}

// x_init <- integrate(x_init, DX_1) in place inside p (one thread per instance); optionally tau_prev <- tau of node 1
// (the compiled-solver branch of the reference loop, run_mpc.py:108-111; the generic branch leaves tau_prev alone)
__global__ void mpc_advance_kernel(const PlmLayout* __restrict__ Lp, int batch, int nq, int nv, int update_tau_prev, const double* __restrict__ x,
                                   double* __restrict__ p) {
  const PlmLayout& L = *Lp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const bool cvel = L.dynamics == PLM_CENTROIDAL_VEL;
  const int qo = cvel ? 6 : 0;
  double* xi = p + (size_t)b * L.np + L.p_x_init;
  const double* d = x + (size_t)b * L.n + L.x_off[1];
  double pos[3], quat[4];
  se3_integrate(xi + qo, xi + qo + 3, d + qo, pos, quat);
  for (int i = 0; i < 3; ++i) xi[qo + i] = pos[i];
  for (int i = 0; i < 4; ++i) xi[qo + 3 + i] = quat[i];
  for (int j = 0; j < nq - 7; ++j) xi[qo + 7 + j] += d[qo + 6 + j];
  if (cvel) { for (int i = 0; i < 6; ++i) xi[i] += d[i]; }
  else { for (int i = 0; i < nv; ++i) xi[nq + i] += d[nv + i]; }
  if (update_tau_prev && L.dynamics == PLM_WHOLE_BODY_RNEA && L.tau_nodes > 1) {
    const double* tau1 = x + (size_t)b * L.n + L.x_off[1] + L.ndx + L.tau_idx;
    double* tp = p + (size_t)b * L.np + L.p_tau_prev;
    for (int j = 0; j < nq - 7; ++j) tp[j] = tau1[j];
  }
}

}  // namespace

extern "C" {

int plm_mpc_step(plm_handle* h, double* d_x, double* d_p, const double* d_t0, double t_add, int32_t gait, double gait_period,
                 const double* dts_host, int32_t x_mode, int32_t update_tau_prev, int32_t batch, double* d_x_new,
                 double* d_stats, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  if (gait < 0 || gait > 2) { h->error = "plm_mpc_step: gait 0 trot, 1 walk, 2 stand"; return 11; }
  if (x_mode < 0 || x_mode > 2) { h->error = "plm_mpc_step: x_mode 0 initial guess, 1 warm start, 2 keep"; return 11; }
  const PlmLayout& L = h->host.layout;
  const PlmModel& M = h->host.model;
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->d_mpc_dts) {
    if (cudaMalloc(&h->d_mpc_dts, PLM_MAXNODES * sizeof(double)) != cudaSuccess) { h->error = "plm_mpc_step: allocation failed"; return 7; }
  }
  if (cudaMemcpyAsync(h->d_mpc_dts, dts_host, (size_t)L.nodes * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) {
    h->error = "plm_mpc_step: copy of the step sizes failed";
    return 7;
  }
  // utils/gait_sequence.py:13-35: contacts and swing period per gait
  const int n_contacts = gait == 0 ? 2 : (gait == 1 ? 3 : 4);
  const double swing_period = gait == 0 ? 0.5 * gait_period : (gait == 1 ? 0.25 * gait_period : gait_period);
  mpc_prepare_kernel<<<batch, 32, 0, s>>>(h->d_layout, batch, gait, gait_period, swing_period, n_contacts, h->d_mpc_dts, d_t0, t_add, M.total_mass,
                                          x_mode, d_x, d_p);
  PLM_LAUNCH_CHECK(h);
  h->launches++;
  if (int rc = plm_sqp_step(h, d_x, d_p, batch, d_x_new, d_stats, stream)) return rc;
  mpc_advance_kernel<<<(batch + 63) / 64, 64, 0, s>>>(h->d_layout, batch, M.nq, M.nv, update_tau_prev, d_x_new, d_p);
  PLM_LAUNCH_CHECK(h);
  h->launches++;
  return 0;
}

}  // extern "C"
