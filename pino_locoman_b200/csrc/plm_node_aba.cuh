// whole_body_aba node: a = ABA(q, v, [0; tau_j], f_ext) and d a / d(dq, dv, tau_j, f).
// Replaces dynamics_whole_body_torque.py:73-103 (aba_dyn) and casadi AD of it.  Computed as
// a = M^-1 (tau - rnea(q, v, 0, f)) with M from the subtree composite inertias (M_cd = J_c . I^C J_d),
// a warp-level Cholesky of M in shared memory, and  d a / d z = -M^-1 d rnea / d z  at the solved a.
#pragma once
#include "plm_node.cuh"

namespace plm {


template <class Exec>
PLM_HD void aba_solve_and_derivatives(Exec& ex, NodeWs& ws, const NodeArgs& A) {
  const PlmModel& M = *A.M;
  const PlmLayout& L = *A.L;
  const int nv = M.nv, nf = L.nf, nj = M.nj;
  // nv x nv scratch matrices with an odd leading dimension: row-per-lane and column-per-lane accesses are both free of
  // bank conflicts.  The Cholesky factor is dead once M^-1 exists, so d rnea / dq reuses its storage.
  const int PLM_LD = nv | 1;
  double* Mi = ws.aba;                 // M^-1
  double* Ms = ws.aba + PLM_LD * nv;   // M, then its Cholesky factor (lower)
  double* GQ = Ms;                     // d rnea / dq (after step 4)
  double* GV = ws.aba + 2 * PLM_LD * nv;
  double* GF = ws.aba + 3 * PLM_LD * nv;   // [nv][nf]
  const double* u = A.xs + L.ndx;

  // 1. bias torques (a = 0), J and I^C J per column
  ex.run([&](int lane, LaneState& st) {
    for (int e = lane; e < PLM_LD * nv; e += 32) Ms[e] = 0.0;
    if (lane >= nv) return;
    const double* rec = ws.rec[M.col_body[lane]];
    st.tau = dot6(st.J, rec + PLM_REC_F);
    inertia_mul(rec[PLM_REC_M], rec + PLM_REC_MC, rec + PLM_REC_IB, st.J, st.w);
    for (int i = 0; i < 6; ++i) ws.col[lane][i] = st.J[i];
  });
  // 2. joint-space inertia
  ex.run([&](int lane, LaneState& st) {
    if (lane >= nv) return;
    const int len = M.chain_len[lane];
    for (int a = 0; a < 6 + len; ++a) {
      const int c = (a < 6) ? a : (M.chain[lane][a - 6] + 5);
      if (M.col_body[lane] == 0 && c > lane) continue;
      double v = dot6(ws.col[c], st.w);
      Ms[c * PLM_LD + lane] = v;
      Ms[lane * PLM_LD + c] = v;
    }
  });
  // 3. Cholesky M = L L^T (right-looking, lane = row; reciprocals on the diagonal)
  for (int k = 0; k < nv; ++k) {
    ex.run([&](int lane, LaneState&) {
      if (lane == k) Ms[k * PLM_LD + k] = 1.0 / sqrt(Ms[k * PLM_LD + k]);      // the diagonal holds 1 / L_kk: no divisions below
    });
    ex.run([&](int lane, LaneState&) {
      if (lane > k && lane < nv) Ms[lane * PLM_LD + k] *= Ms[k * PLM_LD + k];
    });
    ex.run([&](int lane, LaneState&) {
      if (lane > k && lane < nv) {
        const double lik = Ms[lane * PLM_LD + k];
        for (int j = k + 1; j <= lane; ++j) Ms[lane * PLM_LD + j] -= lik * Ms[j * PLM_LD + k];
      }
    });
  }
  // 4. M^-1 column by column (lane = column), then a = M^-1 (tau - bias)
  ex.run([&](int lane, LaneState&) {
    if (lane >= nv) return;
    // forward: L y = e_lane
    for (int i = 0; i < nv; ++i) {
      double s = (i == lane) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= Ms[i * PLM_LD + j] * Mi[j * PLM_LD + lane];
      Mi[i * PLM_LD + lane] = s * Ms[i * PLM_LD + i];
    }
    // backward: L^T x = y
    for (int i = nv - 1; i >= 0; --i) {
      double s = Mi[i * PLM_LD + lane];
      for (int j = i + 1; j < nv; ++j) s -= Ms[j * PLM_LD + i] * Mi[j * PLM_LD + lane];
      Mi[i * PLM_LD + lane] = s * Ms[i * PLM_LD + i];
    }
  });
  ex.run([&](int lane, LaneState& st) {
    if (lane >= nv) return;
    ws.cq[lane] = ((lane >= 6) ? u[lane - 6] : 0.0) - st.tau;   // rhs, staged in cq (restored below)
  });
  ex.run([&](int lane, LaneState&) {
    if (lane >= nv) return;
    double s = 0.0;
    for (int d = 0; d < nv; ++d) s += Mi[d * PLM_LD + lane] * ws.cq[d];
    ws.ca[lane] = s;
  });
  ex.run([&](int lane, LaneState&) {
    if (lane < nv) ws.cq[lane] = A.xs[lane];   // restore dq
  });
  if (!A.want_jac) return;
  // 5. recompute the recursion at the solved acceleration
  ex.run([&](int lane, LaneState& st) { node_phase_b<PLM_WHOLE_BODY_ABA>(ws, A, st, lane); });
  ex.run([&](int lane, LaneState&) { node_phase_c(ws, M, lane); });
  ex.run([&](int lane, LaneState& st) {
    for (int e = lane; e < 2 * PLM_LD * nv + nv * nf; e += 32) GQ[e] = 0.0;   // GQ, GV, GF are contiguous
    node_phase_d<PLM_WHOLE_BODY_ABA>(ws, A, st, lane);
  });
  // 6. d rnea / d(q, v, f) into dense scratch (all related pairs)
  ex.run([&](int lane, LaneState& st) {
    if (lane >= nv) return;
    const int body = M.col_body[lane];
    const unsigned mask = M.col_contacts[lane];
    const int len = M.chain_len[lane];
    for (int a = 0; a < 6 + len; ++a) {
      const int c = (a < 6) ? a : (M.chain[lane][a - 6] + 5);
      const double* cr = ws.col[c];
      const bool same = (M.col_body[c] == body);
      GV[c * PLM_LD + lane] = dot6(cr, st.dFv);
      GQ[c * PLM_LD + lane] = dot6(cr, same ? st.dFqn : st.dFq);
      if (!same) {
        double tv = dot6(st.w, cr + 18) + dot6(st.y, cr);
        double tq = dot6(st.w, cr + 12) + dot6(st.y, cr + 6);
        for (int k = 0; k < M.ncontact; ++k) {
          if (!((mask >> k) & 1u)) continue;
          double g[3], wr[6];
          cross3(c < 6 ? ws.jqb[c] : cr + 3, ws.fx + 3 * k, g);
          point_wrench(ws.con[k], g, wr);
          tq += dot6(st.J, wr);
        }
        GV[lane * PLM_LD + c] = tv;
        GQ[lane * PLM_LD + c] = tq;
      }
    }
    for (int k = 0; k < M.ncontact; ++k) {
      if (!((mask >> k) & 1u)) continue;
      double jk[3];
      cross3(st.J + 3, ws.con[k], jk);
      for (int t = 0; t < 3; ++t) GF[lane * nf + 3 * k + t] = -(st.J[t] + jk[t]);
    }
  });
  // 7. d a / d z = -M^-1 G, scaled by -dt for the rows dv_next - (dv + a dt)
  ex.run([&](int lane, LaneState&) {
    const double dt = A.dt;
    for (int col = lane; col < 2 * nv + nf + nj; col += 32) {
      for (int r = 0; r < nv; ++r) {
        double s = 0.0;
        if (col < nv) {
          for (int k = 0; k < nv; ++k) s += Mi[r * PLM_LD + k] * GQ[k * PLM_LD + col];
          emit(ws, A, PLM_SRC_TQ, r * nv + col, dt * s);
        } else if (col < 2 * nv) {
          const int d = col - nv;
          for (int k = 0; k < nv; ++k) s += Mi[r * PLM_LD + k] * GV[k * PLM_LD + d];
          emit(ws, A, PLM_SRC_TV, r * nv + d, dt * s - ((r == d) ? 1.0 : 0.0));
        } else if (col < 2 * nv + nf) {
          const int d = col - 2 * nv;
          for (int k = 0; k < nv; ++k) s += Mi[r * PLM_LD + k] * GF[k * nf + d];
          emit(ws, A, PLM_SRC_TF, r * nv * 0 + r * nf + d, dt * s);
        } else {
          const int j = col - 2 * nv - nf;
          emit(ws, A, PLM_SRC_TA, r * nv + j, -dt * Mi[r * PLM_LD + 6 + j]);
        }
      }
    }
  });
}

}  // namespace plm
