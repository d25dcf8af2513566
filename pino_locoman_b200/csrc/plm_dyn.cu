// Dynamics* Function entry points (dynamics/*.py of the reference), batched.
// Each call packs its arguments into a two-node probe problem of the matching formulation (x_init = the given
// state, DX = 0, dt = 1), runs the node kernel of that formulation once and unpacks the requested rows /
// Jacobian entries -- the same device code that evaluates the OCP rows, so values and derivatives are identical.
#include <string.h>

#include <vector>

#include "plm_handle.cuh"
#include "plm_vec.cuh"

using namespace plm;

#define DYN_CUDA(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return 10;                                                                       \
    }                                                                                  \
  } while (0)

struct plm_probe {
  plm_handle* h = nullptr;       // probe handle (nodes = 2)
  double *x = nullptr, *p = nullptr, *g = nullptr, *J = nullptr;
  double *rows6 = nullptr, *jac6 = nullptr;   // scratch of the base_vel / base_acc solves
  int32_t* map = nullptr;        // [rows_out][cols_out] -> position in the probe's J values (or -1)
  int rows_out = 0, cols_out = 0;
  int map_kind = 0;              // which entry point built the map (1 value+jac export, 2 base solve)
};

__global__ void probe_pack_kernel(int batch, int n, int np, double* __restrict__ x, double* __restrict__ p,
                                  int p_x_init, int nx_a, const double* __restrict__ a_src, int nx_b, const double* __restrict__ b_src,
                                  int u_off, int nu_a, const double* __restrict__ ua, int nu_b, const double* __restrict__ ub,
                                  int p_dt_min, int p_dt_max, int p_n_contacts, int p_swing_period, int p_contact, int ncontact_flags) {
  const int b = blockIdx.x;
  double* xb = x + (size_t)b * n;
  double* pb = p + (size_t)b * np;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xb[i] = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) pb[i] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < nx_a; i += blockDim.x) pb[p_x_init + i] = a_src ? a_src[(size_t)b * nx_a + i] : 0.0;
  for (int i = threadIdx.x; i < nx_b; i += blockDim.x) pb[p_x_init + nx_a + i] = b_src ? b_src[(size_t)b * nx_b + i] : 0.0;
  for (int i = threadIdx.x; i < nu_a; i += blockDim.x) xb[u_off + i] = ua ? ua[(size_t)b * nu_a + i] : 0.0;
  for (int i = threadIdx.x; i < nu_b; i += blockDim.x) xb[u_off + nu_a + i] = ub ? ub[(size_t)b * nu_b + i] : 0.0;
  for (int i = threadIdx.x; i < ncontact_flags; i += blockDim.x) pb[p_contact + i] = 1.0;   // every foot in contact
  if (threadIdx.x == 0) {
    pb[p_dt_min] = 1.0; pb[p_dt_max] = 1.0; pb[p_n_contacts] = 1.0; pb[p_swing_period] = 1.0;
  }
}

// out[b][r] = sign * g[b][row0 + r]
__global__ void probe_rows_kernel(int batch, int m, int row0, int nrows, double sign, const double* __restrict__ g, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * nrows) return;
  const int b = i / nrows, r = i % nrows;
  out[i] = sign * g[(size_t)b * m + row0 + r];
}

// jac[b][r][c] = sign * J[b][map[r][c]] (+ diag_add on r == c + diag_col0), 0 where the entry is structurally absent
__global__ void probe_jac_kernel(int batch, int nnz, int rows, int cols, const int32_t* __restrict__ map, double sign,
                                 int diag_col0, double diag_add, const double* __restrict__ J, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)batch * rows * cols) return;
  const int b = (int)(i / (rows * cols)), rc = (int)(i % (rows * cols));
  const int r = rc / cols, c = rc % cols;
  const int pos = map[rc];
  double v = pos >= 0 ? sign * J[(size_t)b * nnz + pos] : 0.0;
  if (diag_col0 >= 0 && c == diag_col0 + r) v += diag_add;
  out[i] = v;
}

// x = -A^-1 r for the 6x6 blocks A = jac[b][0:6][col0:col0+6] (row stride ld) and r = rows[b][0:6]; Gaussian elimination
// with partial pivoting, one thread per instance.
__global__ void solve6_kernel(int batch, const double* __restrict__ jac, int ld, int col0, const double* __restrict__ rows, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double A[6][7];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) A[i][j] = jac[((size_t)b * 6 + i) * ld + col0 + j];
    A[i][6] = -rows[(size_t)b * 6 + i];
  }
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    for (int i = k + 1; i < 6; ++i) if (fabs(A[i][k]) > fabs(A[piv][k])) piv = i;
    for (int j = k; j < 7; ++j) { const double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
    const double inv = 1.0 / A[k][k];
    for (int i = k + 1; i < 6; ++i) {
      const double f = A[i][k] * inv;
      for (int j = k; j < 7; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double v = A[i][6];
    for (int j = i + 1; j < 6; ++j) v -= A[i][j] * out[(size_t)b * 6 + j];
    out[(size_t)b * 6 + i] = v / A[i][i];
  }
}

static void probe_free(plm_probe* pr) {
  if (!pr) return;
  cudaFree(pr->x); cudaFree(pr->p); cudaFree(pr->g); cudaFree(pr->J); cudaFree(pr->map); cudaFree(pr->rows6); cudaFree(pr->jac6);
  if (pr->h) plm_destroy(pr->h);
  delete pr;
}

void plm_dyn_free(plm_handle* h) {
  cudaFree(h->d_frame_plc);
  for (int k = 0; k < 5; ++k) { probe_free(h->probes[k]); h->probes[k] = nullptr; }
}

// Lazily create the probe of one formulation; row0/nrows select the rows whose Jacobian is exported, `cols` lists the
// local columns (node-0 block of x) of the exported Jacobian.
static int get_probe(plm_handle* h, int dynamics, plm_probe** out) {
  if (h->is_probe) { h->error = "probe handles do not nest"; return 10; }
  if (h->probes[dynamics]) { *out = h->probes[dynamics]; return 0; }
  const PlmModel& M = h->host.model;
  std::vector<double> placement(12 * M.nbody), axis(3 * M.nbody), inertia(10 * M.nbody), coff(3 * M.ncontact);
  std::vector<int32_t> parent(M.nbody), cbody(M.ncontact);
  for (int b = 0; b < M.nbody; ++b) {
    parent[b] = M.parent[b];
    for (int i = 0; i < 9; ++i) placement[12 * b + i] = M.place_R[b][i];
    for (int i = 0; i < 3; ++i) { placement[12 * b + 9 + i] = M.place_p[b][i]; axis[3 * b + i] = M.axis[b][i]; inertia[10 * b + 1 + i] = M.com[b][i]; }
    inertia[10 * b] = M.mass[b];
    for (int i = 0; i < 6; ++i) inertia[10 * b + 4 + i] = M.Ic[b][i];
  }
  for (int k = 0; k < M.ncontact; ++k) { cbody[k] = M.contact_body[k]; for (int i = 0; i < 3; ++i) coff[3 * k + i] = M.contact_off[k][i]; }
  plm_robot_desc rd;
  memset(&rd, 0, sizeof(rd));
  rd.nbody = M.nbody; rd.parent = parent.data(); rd.placement = placement.data(); rd.axis = axis.data(); rd.inertia = inertia.data();
  rd.nfeet = M.nfeet; rd.has_ext_force = M.has_ext; rd.contact_body = cbody.data(); rd.contact_offset = coff.data();
  rd.arm_body = M.arm_body;
  for (int i = 0; i < 3; ++i) rd.arm_offset[i] = M.arm_off[i];
  rd.joint_pos_min = M.joint_pos_min; rd.joint_pos_max = M.joint_pos_max; rd.joint_vel_max = M.joint_vel_max;
  rd.joint_torque_max = M.joint_torque_max; rd.q0 = M.q0;
  plm_ocp_desc od = h->ocp;
  od.dynamics = dynamics; od.nodes = 2; od.tau_nodes = 2;
  od.include_base = 1; od.include_acc = 1;      // the Dynamics* functions are those of the full formulations
  plm_probe* pr = new plm_probe();
  int rc = plm_create(&rd, &od, h->max_batch, &pr->h);
  if (rc) { h->error = std::string("probe: ") + (pr->h ? pr->h->error : "alloc"); probe_free(pr); return rc; }
  pr->h->is_probe = 1;
  const PlmLayout& L = pr->h->host.layout;
  const size_t B = (size_t)h->max_batch;
  DYN_CUDA(h, cudaMalloc(&pr->x, B * L.n * 8));
  DYN_CUDA(h, cudaMalloc(&pr->p, B * L.np * 8));
  DYN_CUDA(h, cudaMalloc(&pr->g, B * L.m * 8));
  DYN_CUDA(h, cudaMalloc(&pr->J, B * L.nnz * 8));
  h->probes[dynamics] = pr;
  *out = pr;
  return 0;
}

// Build / upload the (row, col) -> J position map for rows [row0, row0 + nrows) of node `node` and the given local columns.
static int set_map(plm_handle* h, plm_probe* pr, int node, int row0, int nrows, const std::vector<int>& cols) {
  const HostTables& T = pr->h->host;
  const PlmLayout& L = T.layout;
  std::vector<int32_t> map((size_t)nrows * cols.size(), -1);
  std::vector<int> colpos(L.n + 1, -1);
  for (size_t c = 0; c < cols.size(); ++c) colpos[L.x_off[node] + cols[c]] = (int)c;
  const int gr0 = L.row_off[node] + row0;
  for (size_t e = 0; e < T.pat_rows.size(); ++e) {
    const int r = T.pat_rows[e] - gr0;
    if (r < 0 || r >= nrows) continue;
    const int c = colpos[T.pat_cols[e]];
    if (c >= 0) map[(size_t)r * cols.size() + c] = (int32_t)e;
  }
  if (pr->map) cudaFree(pr->map);
  DYN_CUDA(h, cudaMalloc(&pr->map, map.size() * sizeof(int32_t)));
  DYN_CUDA(h, cudaMemcpy(pr->map, map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  pr->rows_out = nrows;
  pr->cols_out = (int)cols.size();
  pr->map_kind = 1;
  return 0;
}

static int run_probe(plm_handle* h, plm_probe* pr, int batch, const double* xa, int nxa, const double* xb, int nxb,
                     const double* ua, int nua, const double* ub, int nub, int want_jac, cudaStream_t s) {
  const PlmLayout& L = pr->h->host.layout;
  probe_pack_kernel<<<batch, 128, 0, s>>>(batch, L.n, L.np, pr->x, pr->p, L.p_x_init, nxa, xa, nxb, xb, L.ndx, nua, ua, nub, ub,
                                          L.p_dt_min, L.p_dt_max, L.p_n_contacts, L.p_swing_period, L.p_contact, 4 * L.nodes);
  PLM_LAUNCH_CHECK(h);
  int rc = plm_launch_node_eval(pr->h, pr->x, pr->p, batch, pr->g, pr->J, want_jac, s);
  if (rc) { h->error = pr->h->error; return rc; }
  h->launches++;
  return 0;
}

static int emit_rows(plm_handle* h, plm_probe* pr, int batch, int node, int row0, int nrows, double sign, double* out, cudaStream_t s) {
  const PlmLayout& L = pr->h->host.layout;
  probe_rows_kernel<<<(batch * nrows + 127) / 128, 128, 0, s>>>(batch, L.m, L.row_off[node] + row0, nrows, sign, pr->g, out);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

static int emit_jac(plm_handle* h, plm_probe* pr, int batch, double sign, int diag_col0, double diag_add, double* out, cudaStream_t s) {
  const PlmLayout& L = pr->h->host.layout;
  const long long tot = (long long)batch * pr->rows_out * pr->cols_out;
  probe_jac_kernel<<<(int)((tot + 127) / 128), 128, 0, s>>>(batch, L.nnz, pr->rows_out, pr->cols_out, pr->map, sign, diag_col0, diag_add, pr->J, out);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

static std::vector<int> cols_qvaf(const PlmLayout& L, int nv, int lead) {
  std::vector<int> c;
  for (int i = 0; i < nv; ++i) c.push_back(i);                  // dq (tangent)
  for (int i = 0; i < nv; ++i) c.push_back(nv + i);             // dv
  for (int i = 0; i < lead; ++i) c.push_back(L.ndx + i);        // a (or tau_j)
  for (int i = 0; i < L.nf; ++i) c.push_back(L.ndx + L.f_idx + i);
  return c;
}

extern "C" {

int plm_rnea_dyn(plm_handle* h, const double* d_q, const double* d_v, const double* d_a, const double* d_forces, int32_t batch,
                 double* d_tau, double* d_jac, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  plm_probe* pr;
  if (int rc = get_probe(h, PLM_WHOLE_BODY_RNEA, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  const PlmNodeType& T = L.types[L.node_type[0]];
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = run_probe(h, pr, batch, d_q, M.nq, d_v, M.nv, d_a, M.nv, d_forces, L.nf, d_jac != nullptr, s)) return rc;
  // rows: tau[:6] (row_dyn) and tau[6:] - tau_j (row_tauj, tau_j = 0); they are consecutive
  if (int rc = emit_rows(h, pr, batch, 0, T.row_dyn, M.nv, 1.0, d_tau, s)) return rc;
  if (d_jac) {
    if (pr->map_kind != 1) if (int rc = set_map(h, pr, 0, T.row_dyn, M.nv, cols_qvaf(L, M.nv, M.nv))) return rc;
    if (int rc = emit_jac(h, pr, batch, 1.0, -1, 0.0, d_jac, s)) return rc;
  }
  return 0;
}

int plm_dyn_gaps(plm_handle* h, int32_t dynamics, const double* d_q, const double* d_v, const double* d_a, const double* d_forces,
                 int32_t batch, double* d_gaps, double* d_jac, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  if (dynamics != PLM_CENTROIDAL_ACC && dynamics != PLM_WHOLE_BODY_ACC) { h->error = "plm_dyn_gaps: centroidal_acc or whole_body_acc"; return 11; }
  plm_probe* pr;
  if (int rc = get_probe(h, dynamics, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  const PlmNodeType& T = L.types[L.node_type[0]];
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = run_probe(h, pr, batch, d_q, M.nq, d_v, M.nv, d_a, M.nv, d_forces, L.nf, d_jac != nullptr, s)) return rc;
  if (int rc = emit_rows(h, pr, batch, 0, T.row_dyn, 6, 1.0, d_gaps, s)) return rc;
  if (d_jac) {
    if (pr->map_kind != 1) if (int rc = set_map(h, pr, 0, T.row_dyn, 6, cols_qvaf(L, M.nv, M.nv))) return rc;
    if (int rc = emit_jac(h, pr, batch, 1.0, -1, 0.0, d_jac, s)) return rc;
  }
  return 0;
}

int plm_aba_dyn(plm_handle* h, const double* d_q, const double* d_v, const double* d_tau_j, const double* d_forces, int32_t batch,
                double* d_a, double* d_jac, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  plm_probe* pr;
  if (int rc = get_probe(h, PLM_WHOLE_BODY_ABA, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  const PlmNodeType& T = L.types[L.node_type[0]];
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = run_probe(h, pr, batch, d_q, M.nq, d_v, M.nv, d_tau_j, M.nj, d_forces, L.nf, d_jac != nullptr, s)) return rc;
  // rows dv_next - (dv + a dt) with dv = dv_next = 0, dt = 1  =>  a = -row ; d a / d v = -(entry) - I
  if (int rc = emit_rows(h, pr, batch, 0, T.row_int + M.nv, M.nv, -1.0, d_a, s)) return rc;
  if (d_jac) {
    if (pr->map_kind != 1) if (int rc = set_map(h, pr, 0, T.row_int + M.nv, M.nv, cols_qvaf(L, M.nv, M.nj))) return rc;
    if (int rc = emit_jac(h, pr, batch, -1.0, M.nv, -1.0, d_jac, s)) return rc;
  }
  return 0;
}

int plm_centroidal_vel_gaps(plm_handle* h, const double* d_h, const double* d_q, const double* d_v, int32_t batch, double* d_gaps, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  plm_probe* pr;
  if (int rc = get_probe(h, PLM_CENTROIDAL_VEL, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = run_probe(h, pr, batch, d_h, 6, d_q, M.nq, d_v, M.nv, nullptr, L.nf, 0, s)) return rc;
  return emit_rows(h, pr, batch, 0, L.types[L.node_type[0]].row_dyn, 6, 1.0, d_gaps, s);
}

/* base_acc(q, v, a_j, forces) -> a_b (centroidal_acc / centroidal_vel: dynamics_centroidal_acc.py:43-82; whole_body_acc:
 * dynamics_whole_body_acc.py:43-83) and base_vel(h, q, v_j) -> v_b (dynamics_centroidal_vel.py:73-89).  The path rows are
 * affine in the base part of a (resp. v): rows(lead) = A_b lead_b + rows(lead_b = 0), so lead_b = -A_b^-1 rows(lead_b = 0)
 * with A_b the base block of the analytic Jacobian -- the same 6x6 system the reference inverts symbolically. */
int plm_base_solve(plm_handle* h, int32_t dynamics, const double* d_h, const double* d_q, const double* d_v, const double* d_lead_j,
                   const double* d_forces, int32_t batch, double* d_base, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  if (dynamics != PLM_CENTROIDAL_ACC && dynamics != PLM_WHOLE_BODY_ACC && dynamics != PLM_CENTROIDAL_VEL) { h->error = "plm_base_solve: formulation without a base path constraint"; return 11; }
  plm_probe* pr;
  if (int rc = get_probe(h, dynamics, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  const PlmNodeType& T = L.types[L.node_type[0]];
  cudaStream_t s = (cudaStream_t)stream;
  const bool cvel = dynamics == PLM_CENTROIDAL_VEL;
  if (!pr->rows6) {
    DYN_CUDA(h, cudaMalloc(&pr->rows6, (size_t)h->max_batch * 6 * 8));
    DYN_CUDA(h, cudaMalloc(&pr->jac6, (size_t)h->max_batch * 6 * M.nv * 8));
  }
  // lead = [0 (6) | lead_j]: the probe packs two consecutive blocks, so pass a null first block of size 6
  // (probe_pack writes ua (size 6, null -> zeros) then ub); forces follow only when the lead is exactly nv long, so the
  // forces are packed by a second call-site below for the (a, f) formulations.
  {
    probe_pack_kernel<<<batch, 128, 0, s>>>(batch, L.n, L.np, pr->x, pr->p, L.p_x_init, cvel ? 6 : M.nq, cvel ? d_h : d_q,
                                            cvel ? M.nq : M.nv, cvel ? d_q : d_v, L.ndx, 6, nullptr, M.nj, d_lead_j,
                                            L.p_dt_min, L.p_dt_max, L.p_n_contacts, L.p_swing_period, L.p_contact, 4 * L.nodes);
    PLM_LAUNCH_CHECK(h);
    if (!cvel && d_forces) {
      DYN_CUDA(h, cudaMemcpy2DAsync(pr->x + L.ndx + L.f_idx, (size_t)L.n * 8, d_forces, (size_t)L.nf * 8, (size_t)L.nf * 8, batch,
                                    cudaMemcpyDeviceToDevice, s));
    }
    int rc = plm_launch_node_eval(pr->h, pr->x, pr->p, batch, pr->g, pr->J, 1, s);
    if (rc) { h->error = pr->h->error; return rc; }
    h->launches++;
  }
  if (int rc = emit_rows(h, pr, batch, 0, T.row_dyn, 6, 1.0, pr->rows6, s)) return rc;
  // Jacobian of the 6 path rows w.r.t. the leading input block (a or v): local columns ndx .. ndx + nv
  std::vector<int> cols;
  for (int i = 0; i < M.nv; ++i) cols.push_back(L.ndx + i);
  if (pr->map_kind != 2) {
    if (int rc = set_map(h, pr, 0, T.row_dyn, 6, cols)) return rc;
    pr->map_kind = 2;
  }
  if (int rc = emit_jac(h, pr, batch, 1.0, -1, 0.0, pr->jac6, s)) return rc;
  solve6_kernel<<<(batch + 63) / 64, 64, 0, s>>>(batch, pr->jac6, M.nv, 0, pr->rows6, d_base);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_com_dyn(plm_handle* h, const double* d_q, const double* d_forces, int32_t batch, double* d_dh, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  plm_probe* pr;
  if (int rc = get_probe(h, PLM_CENTROIDAL_VEL, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  cudaStream_t s = (cudaStream_t)stream;
  // x_init = [h = 0 | q], U_0 = [v = 0 | forces]; rows dh_next - (dh + h_dot dt) with dh = dh_next = 0, dt = 1  =>  h_dot = -row
  if (int rc = run_probe(h, pr, batch, nullptr, 6, d_q, M.nq, nullptr, M.nv, d_forces, L.nf, 0, s)) return rc;
  return emit_rows(h, pr, batch, 0, L.types[L.node_type[0]].row_int, 6, -1.0, d_dh, s);
}

int plm_frame_vel(plm_handle* h, int32_t contact, int32_t relative_to_base, const double* d_q, const double* d_v, int32_t batch,
                  double* d_vel, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  plm_probe* pr;
  if (int rc = get_probe(h, PLM_WHOLE_BODY_RNEA, &pr)) return rc;
  const PlmLayout& L = pr->h->host.layout;
  const PlmModel& M = pr->h->host.model;
  const PlmNodeType& T = L.types[L.node_type[1]];      // node 1 carries the state rows
  int row0;
  if (!relative_to_base && contact >= 0 && contact < M.nfeet) row0 = T.row_foot[contact] + 5;
  else if (relative_to_base && contact == -1 && T.row_arm >= 0) row0 = T.row_arm;
  else { h->error = "plm_frame_vel: foot frames (world-aligned) or the arm frame (base-relative) only"; return 11; }
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = run_probe(h, pr, batch, d_q, M.nq, d_v, M.nv, nullptr, M.nv, nullptr, L.nf, 0, s)) return rc;
  return emit_rows(h, pr, batch, 1, row0, 3, 1.0, d_vel, s);
}

__global__ void frame_kin_kernel(const PlmModel* __restrict__ Mp, int batch, int body, const double* __restrict__ plc12, int base_body,
                                 const double* __restrict__ base_plc12, int relative_to_base, const double* __restrict__ q,
                                 const double* __restrict__ v, double* __restrict__ pos, double* __restrict__ vel);

int plm_frame_kinematics(plm_handle* h, int32_t body, const double* placement12, int32_t base_body, const double* base_placement12,
                         int32_t relative_to_base, const double* d_q, const double* d_v, int32_t batch, double* d_pos, double* d_vel,
                         void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmModel& M = h->host.model;
  if (body < 0 || body >= M.nbody || base_body < 0 || base_body >= M.nbody) { h->error = "plm_frame_kinematics: body out of range"; return 11; }
  if (d_vel && !d_v) { h->error = "plm_frame_kinematics: velocities need d_v"; return 11; }
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->d_frame_plc) DYN_CUDA(h, cudaMalloc(&h->d_frame_plc, 24 * sizeof(double)));
  double plc[24];
  for (int i = 0; i < 12; ++i) { plc[i] = placement12[i]; plc[12 + i] = base_placement12 ? base_placement12[i] : (i % 4 == 0 && i < 9 ? 1.0 : 0.0); }
  DYN_CUDA(h, cudaMemcpyAsync(h->d_frame_plc, plc, sizeof(plc), cudaMemcpyHostToDevice, s));
  frame_kin_kernel<<<(batch + 63) / 64, 64, 0, s>>>(h->d_model, batch, body, h->d_frame_plc, base_body, h->d_frame_plc + 12, relative_to_base,
                                                    d_q, d_v, d_pos, d_vel);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

// frame_pos(q) / frame_vel(q, v) of an arbitrary frame (dynamics/dynamics.py:67-118), one thread per instance: forward
// kinematics and the velocity recursion along the chain of the frame's parent body, in world coordinates.
//   pos  = oMf.translation
//   vel  = getFrameVelocity(LOCAL_WORLD_ALIGNED) = [v of the frame origin; omega], world axes; with relative_to_base the
//          reference's variant (dynamics.py:86-113): [R_b^T(v_f - v_b - w_b x (p_f - p_b))_xy, v_f.z, R_b^T(w_f - w_b)_xy, w_f.z]
// Frames are given by their parent body and placement (R row-major | p) in that body's joint frame.
__device__ void frame_state(const PlmModel& M, const double* q, const double* v, int body, const double* plc, double* Rf, double* pf,
                            double* vf, double* wf) {
  double R[9], p[3] = {q[0], q[1], q[2]}, vo[3] = {0, 0, 0}, w[3] = {0, 0, 0};
  quat_to_R(q + 3, R);
  if (v) { matvec3(R, v, vo); matvec3(R, v + 3, w); }      // free-flyer velocity is expressed in the base frame
  if (body > 0) {
    const int col = body + 5, len = M.chain_len[col];
    for (int l = 0; l < len; ++l) {
      const int jb = M.chain[col][l];
      double t[3], t2[3];
      matvec3(R, M.place_p[jb], t);                          // joint origin in world axes, relative to the parent origin
      if (v) { cross3(w, t, t2); for (int i = 0; i < 3; ++i) vo[i] += t2[i]; }
      for (int i = 0; i < 3; ++i) p[i] += t[i];
      if (M.has_rot[jb]) matmul3(R, M.place_R[jb], R);
      double sn, cs;
      sincos(q[7 + jb - 1], &sn, &cs);
      double Rot[9];
      const double x = M.axis[jb][0], y = M.axis[jb][1], z = M.axis[jb][2], tt = 1.0 - cs;
      Rot[0] = cs + tt * x * x;     Rot[1] = tt * x * y - sn * z; Rot[2] = tt * x * z + sn * y;
      Rot[3] = tt * x * y + sn * z; Rot[4] = cs + tt * y * y;     Rot[5] = tt * y * z - sn * x;
      Rot[6] = tt * x * z - sn * y; Rot[7] = tt * y * z + sn * x; Rot[8] = cs + tt * z * z;
      matmul3(R, Rot, R);
      if (v) {
        double ax[3];
        matvec3(R, M.axis[jb], ax);
        const double qd = v[6 + jb - 1];
        for (int i = 0; i < 3; ++i) w[i] += ax[i] * qd;
      }
    }
  }
  double t[3], t2[3];
  matvec3(R, plc + 9, t);
  matmul3(R, plc, Rf);
  for (int i = 0; i < 3; ++i) pf[i] = p[i] + t[i];
  if (v) {
    cross3(w, t, t2);
    for (int i = 0; i < 3; ++i) { vf[i] = vo[i] + t2[i]; wf[i] = w[i]; }
  }
}

__global__ void frame_kin_kernel(const PlmModel* __restrict__ Mp, int batch, int body, const double* __restrict__ plc12, int base_body,
                                 const double* __restrict__ base_plc12, int relative_to_base, const double* __restrict__ q,
                                 const double* __restrict__ v, double* __restrict__ pos, double* __restrict__ vel) {
  const PlmModel& M = *Mp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* qb = q + (size_t)b * M.nq;
  const double* vb = v ? v + (size_t)b * M.nv : nullptr;
  double plc[12], Rf[9], pf[3], vf[3], wf[3];
  for (int i = 0; i < 12; ++i) plc[i] = plc12[i];
  frame_state(M, qb, vb, body, plc, Rf, pf, vf, wf);
  if (pos) for (int i = 0; i < 3; ++i) pos[(size_t)b * 3 + i] = pf[i];
  if (!vel) return;
  double* o = vel + (size_t)b * 6;
  if (!relative_to_base) {
    for (int i = 0; i < 3; ++i) { o[i] = vf[i]; o[3 + i] = wf[i]; }
    return;
  }
  double Rb[9], pb[3], vbf[3], wbf[3], d[3], t[3], lin[3], ang[3], linb[3], angb[3];
  for (int i = 0; i < 12; ++i) plc[i] = base_plc12[i];
  frame_state(M, qb, vb, base_body, plc, Rb, pb, vbf, wbf);
  for (int i = 0; i < 3; ++i) d[i] = pf[i] - pb[i];
  cross3(wbf, d, t);
  for (int i = 0; i < 3; ++i) { lin[i] = vf[i] - vbf[i] - t[i]; ang[i] = wf[i] - wbf[i]; }
  matTvec3(Rb, lin, linb);
  matTvec3(Rb, ang, angb);
  o[0] = linb[0]; o[1] = linb[1]; o[2] = vf[2];      // z components stay in the world frame (dynamics.py:108-113)
  o[3] = angb[0]; o[4] = angb[1]; o[5] = wf[2];
}

__global__ void state_integrate_kernel(int batch, int nq, int nv, int cvel, const double* __restrict__ x, const double* __restrict__ dx,
                                       double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int nx = cvel ? 6 + nq : nq + nv, ndx = cvel ? 6 + nv : 2 * nv, qo = cvel ? 6 : 0;
  const double* xb = x + (size_t)b * nx;
  const double* db = dx + (size_t)b * ndx;
  double* ob = out + (size_t)b * nx;
  se3_integrate(xb + qo, xb + qo + 3, db + qo, ob + qo, ob + qo + 3);
  for (int j = 0; j < nq - 7; ++j) ob[qo + 7 + j] = xb[qo + 7 + j] + db[qo + 6 + j];
  if (cvel) { for (int i = 0; i < 6; ++i) ob[i] = xb[i] + db[i]; }
  else { for (int i = 0; i < nv; ++i) ob[nq + i] = xb[nq + i] + db[nv + i]; }
}

__global__ void state_difference_kernel(int batch, int nq, int nv, int cvel, const double* __restrict__ x0, const double* __restrict__ x1,
                                        double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int nx = cvel ? 6 + nq : nq + nv, ndx = cvel ? 6 + nv : 2 * nv, qo = cvel ? 6 : 0;
  const double* a = x0 + (size_t)b * nx;
  const double* c = x1 + (size_t)b * nx;
  double* ob = out + (size_t)b * ndx;
  se3_difference(a + qo, a + qo + 3, c + qo, c + qo + 3, ob + qo);
  for (int j = 0; j < nq - 7; ++j) ob[qo + 6 + j] = c[qo + 7 + j] - a[qo + 7 + j];
  if (cvel) { for (int i = 0; i < 6; ++i) ob[i] = c[i] - a[i]; }
  else { for (int i = 0; i < nv; ++i) ob[nv + i] = c[nq + i] - a[nq + i]; }
}

int plm_state_integrate(plm_handle* h, const double* d_x, const double* d_dx, int32_t batch, double* d_x_next, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmModel& M = h->host.model;
  state_integrate_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(batch, M.nq, M.nv, h->host.layout.dynamics == PLM_CENTROIDAL_VEL, d_x, d_dx, d_x_next);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_state_difference(plm_handle* h, const double* d_x0, const double* d_x1, int32_t batch, double* d_dx, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmModel& M = h->host.model;
  state_difference_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(batch, M.nq, M.nv, h->host.layout.dynamics == PLM_CENTROIDAL_VEL, d_x0, d_x1, d_dx);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

}  // extern "C"
