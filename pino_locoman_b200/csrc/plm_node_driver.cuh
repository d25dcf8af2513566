// Phase sequencing of one node evaluation, shared by the CUDA kernel (one warp, phases separated by
// __syncwarp) and by the host emulation used in the CPU tests (lanes of a phase run in a loop).
#pragma once
#include "plm_node.cuh"

namespace plm {

PLM_HD size_t aba_ws_doubles(int nv, int nf) { return (size_t)3 * (nv | 1) * nv + (size_t)nv * nf; }   // M^-1, L | dq block, dv block, df block

#if defined(__CUDACC__)
struct WarpExec {
  int lane;
  LaneState st;
  template <class F>
  __device__ __forceinline__ void run(F f) {
    f(lane, st);
    __syncwarp();
  }
};
#endif

struct HostExec {
  LaneState st[32];
  template <class F>
  void run(F f) {
    for (int l = 0; l < 32; ++l) f(l, st[l]);
  }
};

// Bytes of shared workspace one warp needs for a layout.
// Doubles of workspace one warp needs.  stage_J: keep a staging copy of the node's J block in the workspace
// (host emulation); the kernel writes J entries straight into the instance's block in HBM instead.
// with_xbuf: room for the staged trial point x + alpha dx (line-search launches only: evaluation launches run without it,
// which is what lets a 16th warp share the SM).
PLM_HD size_t node_ws_doubles(const PlmLayout& L, int nv, int nf, int nbody, bool stage_J, bool with_xbuf = true) {
  size_t base = (sizeof(NodeWs) + 7) / 8;
  size_t extra = (size_t)nbody * PLM_REC + (size_t)nv * PLM_COLREC + (size_t)L.max_rows + (stage_J ? (size_t)L.max_nnz : 0) +
                 (with_xbuf ? (size_t)(2 * L.ndx + (L.x_off[1] - L.x_off[0] - L.ndx)) : 0);
  if (L.dynamics == PLM_WHOLE_BODY_ABA) extra += aba_ws_doubles(nv, nf);
  if (L.nobase) extra += PLM_VB_DOUBLES;
  return base + extra + 2;
}

// Records, row staging, trial staging and ABA scratch are carved from `tail` (doubles following the NodeWs struct).
PLM_HD void node_ws_bind(NodeWs& ws, const PlmLayout& L, int nv, int nbody, double* tail, double* J_external, bool with_xbuf = true) {
  ws.rec = reinterpret_cast<double (*)[PLM_REC]>(tail);
  tail += (size_t)nbody * PLM_REC;
  ws.col = reinterpret_cast<double (*)[PLM_COLREC]>(tail);
  tail += (size_t)nv * PLM_COLREC;
  ws.g = tail;
  tail += L.max_rows;
  if (J_external) ws.J = J_external;
  else { ws.J = tail; tail += L.max_nnz; }
  ws.xbuf = tail;
  ws.aba = ws.xbuf + (with_xbuf ? (2 * L.ndx + (L.x_off[1] - L.x_off[0] - L.ndx)) : 0);
  ws.vb = ws.aba;      // (the ABA formulation has no variant without base inputs: the two scratch areas never coexist)
}

}  // namespace plm
#include "plm_node_aba.cuh"
namespace plm {

template <int KIND, bool NOBASE, class Exec>
PLM_HD void node_eval_body(Exec& ex, NodeWs& ws, const NodeArgs& A) {
  const PlmModel& M = *A.M;
  const PlmNodeType& T = *A.T;
  (void)T;
  ex.run([&](int lane, LaneState&) { node_phase_a<KIND>(ws, A, lane); });
  ex.run([&](int lane, LaneState& st) { node_phase_b<KIND>(ws, A, st, lane); });
  ex.run([&](int lane, LaneState&) { node_phase_c(ws, M, lane); });
  if (NOBASE) {
    // inputs without the base part: solve the six gap rows for it (first pass ran with a zero base part), then
    // repeat the chain walk and the composites with it in place
    ex.run([&](int lane, LaneState& st) { node_phase_base_cols<KIND>(ws, A, st, lane); });
    ex.run([&](int lane, LaneState&) { node_phase_base_solve<KIND>(ws, lane); });
    ex.run([&](int lane, LaneState& st) { node_phase_b<KIND>(ws, A, st, lane); });
    ex.run([&](int lane, LaneState&) { node_phase_c(ws, M, lane); });
  }
  if (KIND == PLM_WHOLE_BODY_ABA) {
    aba_solve_and_derivatives<Exec>(ex, ws, A);
    ex.run([&](int lane, LaneState& st) {
      node_phase_f<KIND, false>(ws, A, st, lane);
      if (A.want_jac) node_phase_consts(ws, A, lane, 32);
    });
  } else {
    // contact / arm / shared rows first: they only need the chain state of phase B, which keeps the register
    // footprint of the derivative phases (D, E) down
    ex.run([&](int lane, LaneState& st) {
      node_phase_f<KIND, NOBASE>(ws, A, st, lane);
      if (A.want_jac) node_phase_consts(ws, A, lane, 32);
      node_phase_d<KIND>(ws, A, st, lane);
    });
    ex.run([&](int lane, LaneState& st) { node_phase_e<KIND, NOBASE>(ws, A, st, lane); });
    if (NOBASE && A.want_jac) {
      if (KIND == PLM_WHOLE_BODY_ACC)      // force columns of the base rows: written row-wise by lanes 0..5 in phase E
        ex.run([&](int lane, LaneState&) {
          if (lane < A.L->nf) emit_base_rows(ws, A, PLM_SRC_TF, A.L->nf, lane, ws.vb + PLM_VB_GF + 6 * lane, false);
        });
      if (KIND == PLM_CENTROIDAL_VEL)
        ex.run([&](int lane, LaneState& st) { node_phase_fjac<KIND, 2>(ws, A, st, lane); });
    }
  }
}

}  // namespace plm
