// Host emulation of the node-evaluation kernel for CPU tests: the device code of
// pino_locoman_b200/csrc/plm_node*.cuh compiled with g++, the 32 lanes of a warp run in a loop per phase.
// TEST INFRASTRUCTURE: validates the kernel mathematics against the oracle without a GPU; it is not a product path
// (the product library refuses to run without a CUDA device).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../pino_locoman_b200/csrc/plm_host.h"
#include "../../pino_locoman_b200/csrc/plm_node_driver.cuh"

using namespace plm;

struct Emu {
  HostTables t;
};

template <int KIND>
static void run_node(const HostTables& t, const double* x, const double* p, int b, int node, double* g, double* Jv, int want_jac) {
  const PlmLayout& L = t.layout;
  // poison the staging area: an entry of the pattern that no lane writes shows up as NaN in the tests
  std::vector<double> buf(node_ws_doubles(L, t.model.nv, L.nf, t.model.nbody, true) + 8, nan(""));
  NodeWs& ws = *reinterpret_cast<NodeWs*>(buf.data());
  node_ws_bind(ws, L, t.model.nv, t.model.nbody, buf.data() + (sizeof(NodeWs) + 7) / 8, nullptr);
  NodeArgs A;
  A.M = &t.model;
  A.L = &L;
  A.T = &L.types[L.node_type[node]];
  A.lut = t.lut.data() + A.T->lut_off;
  A.consts = t.consts.data() + A.T->const_off;
  A.xs = x + (size_t)b * L.n + L.x_off[node];
  A.p = p + (size_t)b * L.np;
  A.node = node;
  A.dt = node_dt(L, A.p, node);
  A.want_jac = want_jac;
  HostExec ex;
  memset(&ex, 0, sizeof(ex));
  constexpr bool has_variant = KIND == PLM_CENTROIDAL_VEL || KIND == PLM_CENTROIDAL_ACC || KIND == PLM_WHOLE_BODY_ACC;
  if (has_variant && t.layout.nobase) node_eval_body<KIND, has_variant>(ex, ws, A);      // include_base = False
  else node_eval_body<KIND, false>(ex, ws, A);
  const PlmNodeType& T = *A.T;
  for (int r = 0; r < T.nrows; ++r) g[(size_t)b * L.m + L.row_off[node] + r] = ws.g[r];
  if (node == 0) for (int r = 0; r < L.ndx; ++r) g[(size_t)b * L.m + r] = A.xs[r];
  if (want_jac) {
    for (int e = 0; e < T.nnz; ++e) Jv[(size_t)b * L.nnz + L.nnz_off[node] + e] = ws.J[e];
    if (node == 0) for (int e = 0; e < L.ndx; ++e) Jv[(size_t)b * L.nnz + e] = 1.0;
  }
}

extern "C" {

Emu* emu_create(const plm_robot_desc* r, const plm_ocp_desc* o, char* err, int errlen) {
  Emu* e = new Emu();
  if (!build_tables(*r, *o, e->t)) {
    strncpy(err, e->t.error.c_str(), errlen - 1);
    delete e;
    return nullptr;
  }
  return e;
}
void emu_destroy(Emu* e) { delete e; }
void emu_dims(const Emu* e, int* out) {
  const PlmLayout& L = e->t.layout;
  int v[8] = {L.n, L.m, L.np, L.nnz, L.ndx, L.nx, L.nf, L.nodes};
  memcpy(out, v, sizeof(v));
}
void emu_pattern(const Emu* e, int* rows, int* cols) {
  memcpy(rows, e->t.pat_rows.data(), e->t.pat_rows.size() * sizeof(int));
  memcpy(cols, e->t.pat_cols.data(), e->t.pat_cols.size() * sizeof(int));
}
void emu_node_eval(const Emu* e, const double* x, const double* p, int batch, double* g, double* Jv, int want_jac) {
  const PlmLayout& L = e->t.layout;
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < L.nodes; ++i) switch (L.dynamics) {
        case PLM_CENTROIDAL_VEL: run_node<PLM_CENTROIDAL_VEL>(e->t, x, p, b, i, g, Jv, want_jac); break;
        case PLM_CENTROIDAL_ACC: run_node<PLM_CENTROIDAL_ACC>(e->t, x, p, b, i, g, Jv, want_jac); break;
        case PLM_WHOLE_BODY_ACC: run_node<PLM_WHOLE_BODY_ACC>(e->t, x, p, b, i, g, Jv, want_jac); break;
        case PLM_WHOLE_BODY_ABA: run_node<PLM_WHOLE_BODY_ABA>(e->t, x, p, b, i, g, Jv, want_jac); break;
        default: run_node<PLM_WHOLE_BODY_RNEA>(e->t, x, p, b, i, g, Jv, want_jac); break;
      }
}
// storage of a symmetric inverse stage block by cyclic diagonals (plm_qp_types.h): the index arithmetic shared by the
// factor kernel (writer) and the ADMM kernel (reader), and the forward panels of the host-built schedule
int emu_sinv_rows(int s) { return plm_sinv_rows(s); }
int emu_sinv_index(int s, int r, int c) { return plm_sinv_index(s, r, c); }
int emu_qp_schedule(const Emu* e, int latency, int* out, int cap) {
  const QpLayout& Q = e->t.qp;
  const int n = latency ? Q.n_sched_lat : Q.n_sched, off = latency ? Q.f_sched_lat : Q.f_sched;
  for (int k = 0; k < n * PLM_SCHED_INTS && k < cap; ++k) out[k] = e->t.qp_idx32[off + k];
  return n;
}
void emu_qp_factor_offsets(const Emu* e, int* fac_off, int* bk_off, int* panel_doubles) {
  const QpLayout& Q = e->t.qp;
  const int N = e->t.layout.nodes;
  for (int i = 0; i <= N + 1; ++i) fac_off[i] = Q.fac_off[i];
  for (int i = 0; i <= N; ++i) bk_off[i] = Q.bk_off[i];
  panel_doubles[0] = Q.panel_doubles; panel_doubles[1] = Q.panel_doubles_lat; panel_doubles[2] = Q.fac_total;
}
}
