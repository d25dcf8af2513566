// CasADi-external-compatible single-instance entry points.
//
// The reference can swap its SX Functions for compiled ones with
//     self.sqp_data = ca.external("sqp_data", "<lib>.so")          (optimization/ocp.py:299-301)
// This library exports the symbols casadi's external-function loader resolves for the four data Functions of
// optimization/ocp.py:287-290 -- sqp_data(x,p)->[grad_f,J_g,g,lbg,ubg], hess_data(x,p)->[hess_f],
// f_data(x,p)->[f,grad_f], g_data(x,p)->[g,lbg,ubg] -- and evaluates them with libpinolocoman_b200 (batch of one,
// host buffers in / out).  The problem (robot tables + formulation) comes from a file written by
// pino_locoman_b200.casadi_shim.export_problem(); its path is read from the environment variable PLM_CASADI_PROBLEM.
// Conventions (casadi C API): casadi_int = long long, casadi_real = double, sparsity in compressed column storage
// [nrow, ncol, colind[ncol+1], row[nnz]], sparse outputs as their nonzeros in that order, NULL arg = zeros,
// NULL res = not requested.  Not thread safe (one problem, one work area).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pino_locoman_b200.h"

typedef long long casadi_int;
typedef double casadi_real;

namespace {

struct Shim {
  bool ok = false;
  std::string error;
  plm_handle* layout = nullptr;    // layout-only handle: dims and pattern without a GPU
  plm_handle* h = nullptr;         // compute handle (batch 1), created on the first evaluation
  plm_dims d{};
  // problem file image
  std::vector<int32_t> parent, contact_body;
  std::vector<double> placement, axis, inertia, contact_offset, jmin, jmax, vmax, tmax, q0;
  plm_robot_desc rd{};
  plm_ocp_desc od{};
  // sparsities
  std::vector<casadi_int> sp_x, sp_p, sp_g, sp_J, sp_H, sp_1;
  std::vector<int32_t> ccs_src;    // CCS position -> position in the library's (CSR) value order
  // device / host work
  double *dx = nullptr, *dp = nullptr, *dgrad = nullptr, *dJ = nullptr, *dg = nullptr, *dl = nullptr, *du = nullptr, *df = nullptr;
  std::vector<double> hJ, zeros;
};

Shim S;

std::vector<casadi_int> dense_sp(int nrow, int ncol) {
  std::vector<casadi_int> sp = {nrow, ncol};
  for (int c = 0; c <= ncol; ++c) sp.push_back((casadi_int)c * nrow);
  for (int c = 0; c < ncol; ++c) for (int r = 0; r < nrow; ++r) sp.push_back(r);
  return sp;
}

template <class T> bool rd_vec(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

bool load_problem() {
  if (S.ok) return true;
  const char* path = getenv("PLM_CASADI_PROBLEM");
  if (!path) { S.error = "PLM_CASADI_PROBLEM is not set"; return false; }
  FILE* f = fopen(path, "rb");
  if (!f) { S.error = std::string("cannot open ") + path; return false; }
  int32_t hdr[10];
  double arm_off[3], mu;
  bool good = fread(hdr, sizeof(int32_t), 10, f) == 10 && hdr[0] == 0x504d4c50 /* "PLMP" */ && hdr[1] == 1;
  // hdr: magic, version, dynamics, nodes, tau_nodes, nbody, nfeet, has_ext_force, arm_body, nq
  if (good) good = fread(&mu, sizeof(double), 1, f) == 1 && fread(arm_off, sizeof(double), 3, f) == 3;
  const int nbody = good ? hdr[5] : 0, ncontact = good ? hdr[6] + hdr[7] : 0, nj = nbody - 1, nq = good ? hdr[9] : 0;
  good = good && rd_vec(f, S.parent, nbody) && rd_vec(f, S.contact_body, ncontact) && rd_vec(f, S.placement, 12 * (size_t)nbody) &&
         rd_vec(f, S.axis, 3 * (size_t)nbody) && rd_vec(f, S.inertia, 10 * (size_t)nbody) && rd_vec(f, S.contact_offset, 3 * (size_t)ncontact) &&
         rd_vec(f, S.jmin, nj) && rd_vec(f, S.jmax, nj) && rd_vec(f, S.vmax, nj) && rd_vec(f, S.tmax, nj) && rd_vec(f, S.q0, nq);
  fclose(f);
  if (!good) { S.error = std::string("malformed problem file ") + path; return false; }
  plm_robot_desc& r = S.rd;
  r.nbody = nbody; r.parent = S.parent.data(); r.placement = S.placement.data(); r.axis = S.axis.data(); r.inertia = S.inertia.data();
  r.nfeet = hdr[6]; r.has_ext_force = hdr[7]; r.contact_body = S.contact_body.data(); r.contact_offset = S.contact_offset.data();
  r.arm_body = hdr[8];
  for (int i = 0; i < 3; ++i) r.arm_offset[i] = arm_off[i];
  r.joint_pos_min = S.jmin.data(); r.joint_pos_max = S.jmax.data(); r.joint_vel_max = S.vmax.data(); r.joint_torque_max = S.tmax.data();
  r.q0 = S.q0.data();
  plm_fill_default_ocp_desc(&S.od, hdr[2], hdr[3]);
  S.od.tau_nodes = hdr[4];
  S.od.mu = mu;
  if (plm_create(&S.rd, &S.od, 0, &S.layout)) { S.error = S.layout ? plm_last_error(S.layout) : "plm_create failed"; return false; }
  plm_get_dims(S.layout, &S.d);
  const int n = S.d.n, m = S.d.m, nnz = S.d.nnz;
  S.sp_x = dense_sp(n, 1); S.sp_p = dense_sp(S.d.np, 1); S.sp_g = dense_sp(m, 1); S.sp_1 = dense_sp(1, 1);
  // J_g: the library's pattern is COO in CSR value order; casadi wants CCS (column major, rows ascending in a column)
  std::vector<int32_t> rows(nnz), cols(nnz);
  plm_jac_pattern(S.layout, rows.data(), cols.data());
  S.ccs_src.resize(nnz);
  for (int e = 0; e < nnz; ++e) S.ccs_src[e] = e;
  std::stable_sort(S.ccs_src.begin(), S.ccs_src.end(), [&](int a, int b) { return cols[a] != cols[b] ? cols[a] < cols[b] : rows[a] < rows[b]; });
  S.sp_J = {m, n};
  std::vector<casadi_int> colind(n + 1, 0);
  for (int e = 0; e < nnz; ++e) colind[cols[e] + 1]++;
  for (int c = 0; c < n; ++c) colind[c + 1] += colind[c];
  S.sp_J.insert(S.sp_J.end(), colind.begin(), colind.end());
  for (int e = 0; e < nnz; ++e) S.sp_J.push_back(rows[S.ccs_src[e]]);
  // hess_f: diagonal (optimization/ocp.py:293-296)
  S.sp_H = {n, n};
  for (int c = 0; c <= n; ++c) S.sp_H.push_back(c);
  for (int c = 0; c < n; ++c) S.sp_H.push_back(c);
  S.hJ.resize(nnz);
  S.zeros.assign(std::max(n, S.d.np), 0.0);
  S.ok = true;
  return true;
}

bool ensure_compute() {
  if (S.h) return true;
  if (!load_problem()) return false;
  if (plm_create(&S.rd, &S.od, 1, &S.h)) { S.error = S.h ? plm_last_error(S.h) : "plm_create failed (no CUDA device?)"; S.h = nullptr; return false; }
  const plm_dims& d = S.d;
  bool a = cudaMalloc(&S.dx, d.n * 8) == cudaSuccess && cudaMalloc(&S.dp, d.np * 8) == cudaSuccess && cudaMalloc(&S.dgrad, d.n * 8) == cudaSuccess &&
           cudaMalloc(&S.dJ, (size_t)d.nnz * 8) == cudaSuccess && cudaMalloc(&S.dg, d.m * 8) == cudaSuccess && cudaMalloc(&S.dl, d.m * 8) == cudaSuccess &&
           cudaMalloc(&S.du, d.m * 8) == cudaSuccess && cudaMalloc(&S.df, 8) == cudaSuccess;
  if (!a) { S.error = "device allocation failed"; return false; }
  return true;
}

bool upload(const casadi_real** arg) {
  const plm_dims& d = S.d;
  return cudaMemcpy(S.dx, arg[0] ? arg[0] : S.zeros.data(), d.n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(S.dp, arg[1] ? arg[1] : S.zeros.data(), d.np * 8, cudaMemcpyHostToDevice) == cudaSuccess;
}
bool download(double* dst, const double* src, size_t count) { return !dst || cudaMemcpy(dst, src, count * 8, cudaMemcpyDeviceToHost) == cudaSuccess; }

int fail(const char* what) {
  fprintf(stderr, "pino_locoman_b200 casadi shim: %s: %s\n", what, S.error.c_str());
  return 1;
}

const char* kIn[2] = {"x", "p"};

}  // namespace

#define PLM_SHIM_COMMON(NAME, NOUT, WORK_RES)                                                                     \
  extern "C" casadi_int NAME##_n_in(void) { return 2; }                                                             \
  extern "C" casadi_int NAME##_n_out(void) { return NOUT; }                                                         \
  extern "C" const char* NAME##_name_in(casadi_int i) { return i >= 0 && i < 2 ? kIn[i] : nullptr; }                \
  extern "C" const casadi_int* NAME##_sparsity_in(casadi_int i) {                                                   \
    if (!load_problem()) { fail(#NAME "_sparsity_in"); return nullptr; }                                            \
    return i == 0 ? S.sp_x.data() : (i == 1 ? S.sp_p.data() : nullptr);                                             \
  }                                                                                                                 \
  extern "C" int NAME##_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w) {          \
    if (sz_arg) *sz_arg = 2;                                                                                        \
    if (sz_res) *sz_res = WORK_RES;                                                                                 \
    if (sz_iw) *sz_iw = 0;                                                                                          \
    if (sz_w) *sz_w = 0;                                                                                            \
    return 0;                                                                                                       \
  }                                                                                                                 \
  extern "C" void NAME##_incref(void) {}                                                                            \
  extern "C" void NAME##_decref(void) {}

PLM_SHIM_COMMON(sqp_data, 5, 5)
PLM_SHIM_COMMON(hess_data, 1, 1)
PLM_SHIM_COMMON(f_data, 2, 2)
PLM_SHIM_COMMON(g_data, 3, 3)

extern "C" {

const char* sqp_data_name_out(casadi_int i) { static const char* n[5] = {"grad_f", "J_g", "g", "lbg", "ubg"}; return i >= 0 && i < 5 ? n[i] : nullptr; }
const char* hess_data_name_out(casadi_int i) { return i == 0 ? "hess_f" : nullptr; }
const char* f_data_name_out(casadi_int i) { static const char* n[2] = {"f", "grad_f"}; return i >= 0 && i < 2 ? n[i] : nullptr; }
const char* g_data_name_out(casadi_int i) { static const char* n[3] = {"g", "lbg", "ubg"}; return i >= 0 && i < 3 ? n[i] : nullptr; }

const casadi_int* sqp_data_sparsity_out(casadi_int i) {
  if (!load_problem()) { fail("sqp_data_sparsity_out"); return nullptr; }
  switch (i) { case 0: return S.sp_x.data(); case 1: return S.sp_J.data(); case 2: case 3: case 4: return S.sp_g.data(); default: return nullptr; }
}
const casadi_int* hess_data_sparsity_out(casadi_int i) {
  if (!load_problem()) { fail("hess_data_sparsity_out"); return nullptr; }
  return i == 0 ? S.sp_H.data() : nullptr;
}
const casadi_int* f_data_sparsity_out(casadi_int i) {
  if (!load_problem()) { fail("f_data_sparsity_out"); return nullptr; }
  return i == 0 ? S.sp_1.data() : (i == 1 ? S.sp_x.data() : nullptr);
}
const casadi_int* g_data_sparsity_out(casadi_int i) {
  if (!load_problem()) { fail("g_data_sparsity_out"); return nullptr; }
  return i >= 0 && i < 3 ? S.sp_g.data() : nullptr;
}

int sqp_data(const casadi_real** arg, casadi_real** res, casadi_int*, casadi_real*, int) {
  if (!ensure_compute() || !upload(arg)) return fail("sqp_data");
  if (plm_sqp_data(S.h, S.dx, S.dp, 1, S.dgrad, S.dJ, S.dg, S.dl, S.du, nullptr)) { S.error = plm_last_error(S.h); return fail("sqp_data"); }
  const plm_dims& d = S.d;
  bool ok = download(res[0], S.dgrad, d.n) && download(res[2], S.dg, d.m) && download(res[3], S.dl, d.m) && download(res[4], S.du, d.m);
  if (ok && res[1]) {
    ok = download(S.hJ.data(), S.dJ, d.nnz);
    for (int e = 0; ok && e < d.nnz; ++e) res[1][e] = S.hJ[S.ccs_src[e]];
  }
  if (!ok) { S.error = cudaGetErrorString(cudaGetLastError()); return fail("sqp_data"); }
  return 0;
}

int hess_data(const casadi_real** arg, casadi_real** res, casadi_int*, casadi_real*, int) {
  if (!ensure_compute() || !upload(arg)) return fail("hess_data");
  if (plm_hess_diag(S.h, S.dp, 1, S.dgrad, nullptr)) { S.error = plm_last_error(S.h); return fail("hess_data"); }
  if (!download(res[0], S.dgrad, S.d.n)) { S.error = cudaGetErrorString(cudaGetLastError()); return fail("hess_data"); }
  return 0;
}

int f_data(const casadi_real** arg, casadi_real** res, casadi_int*, casadi_real*, int) {
  if (!ensure_compute() || !upload(arg)) return fail("f_data");
  if (plm_f_data(S.h, S.dx, S.dp, 1, S.df, S.dgrad, nullptr)) { S.error = plm_last_error(S.h); return fail("f_data"); }
  if (!(download(res[0], S.df, 1) && download(res[1], S.dgrad, S.d.n))) { S.error = cudaGetErrorString(cudaGetLastError()); return fail("f_data"); }
  return 0;
}

int g_data(const casadi_real** arg, casadi_real** res, casadi_int*, casadi_real*, int) {
  if (!ensure_compute() || !upload(arg)) return fail("g_data");
  if (plm_g_data(S.h, S.dx, S.dp, 1, S.dg, S.dl, S.du, nullptr)) { S.error = plm_last_error(S.h); return fail("g_data"); }
  const plm_dims& d = S.d;
  if (!(download(res[0], S.dg, d.m) && download(res[1], S.dl, d.m) && download(res[2], S.du, d.m))) {
    S.error = cudaGetErrorString(cudaGetLastError());
    return fail("g_data");
  }
  return 0;
}

}  // extern "C"
