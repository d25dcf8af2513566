// Per-shooting-node evaluation of the path-constraint rows and their analytic Jacobian block.
//
// One warp evaluates one (instance, node); lane d owns velocity column d of the kinematic tree.
// The recursion is written in world-aligned spatial coordinates about the base origin (all rows are
// invariant to base translation), so that every lane can walk its own root->joint chain without
// cross-lane traffic, and the subtree composites (net force, inertia, Coriolis factor) are accumulated
// in the warp's shared-memory workspace.  The phases below are separated by warp barriers in the
// kernel; tests/ compiles the same code for the host and runs the lanes of each phase in a loop.
//
// Replaces (reference file:line):
//   dynamics_whole_body_torque.py:42-71  rnea_dyn (+ casadi AD of it, optimization/ocp.py:283)
//   dynamics_whole_body_acc.py:85-126    dyn_gaps = rnea[:6]
//   dynamics_centroidal_acc.py:84-119    dyn_gaps = A a + Adot v - dh
//   dynamics_centroidal_vel.py:43-71,136-148  com_dyn, dyn_gaps
//   dynamics/dynamics.py:77-118          frame_vel (LOCAL_WORLD_ALIGNED, base-relative variant)
//   dynamics_*.py state_integrate        q = integrate(q_init, dq)  (base: R = R0 Exp(w))
//   optimization/ocp.py:103-198 + ocp_*.py setup_dynamics_constraints  row order and canonical forms
#pragma once
#include "plm_types.h"
#include "plm_vec.cuh"

namespace plm {

// body record: net force F(6) | m | mc(3) | Ib(6) | momentum H(6) | B22(9)
#define PLM_REC 31
#define PLM_REC_F 0
#define PLM_REC_M 6
#define PLM_REC_MC 7
#define PLM_REC_IB 10
#define PLM_REC_H 16
#define PLM_REC_B 22
// column record: J(6) | phi_q(6) | chi_q(6) | psi(6)   (the angular q-direction of a joint column is J[3..5]; the
// six base columns keep theirs in NodeWs::jqb)
#define PLM_COLREC 24
// contact record: p_k(3) | body velocity(6)
#define PLM_CONREC 9

struct NodeWs {
  double sc[PLM_MAXB][2];
  double cq[PLM_MAXCOL];      // dq (tangent increment of this node) per column
  double cv[PLM_MAXCOL];      // velocity coordinate per column
  double ca[PLM_MAXCOL];      // acceleration coordinate per column
  double (*rec)[PLM_REC];     // [nbody] body records (carved from the tail of the workspace)
  double (*col)[PLM_COLREC];  // [nv]    column records
  double con[PLM_MAXC][PLM_CONREC];
  double arm[PLM_CONREC];
  double Rb[9];
  double Rinit[9];
  double jr[9];
  double jqb[6][3];           // angular q-direction of the base columns (right Jacobian of Exp mixes them)
  double fx[3 * PLM_MAXC];    // contact / external forces of this node (staged once: every phase reads them)
  double* g;     // [max_rows]
  double* J;     // node block of the Jacobian values: the instance's block in HBM (kernel) or a staging array (host emulation)
  double* aba;   // ABA scratch: M/L, Minv, GQ, GV (nv x 32 each), GF (nv x nf)
  double* xbuf;  // [2 ndx + nu] staged x + alpha dx of this node (line-search trials)
  double* vb;    // formulations without base inputs (PLM_VB_DOUBLES): A_b^-1 (36) | gaps at zero base part -> base part (6) |
                 // d(foot / arm velocity rows)/d v_b (15 x 6) | force columns of the base rows (18 x 6)
};
#define PLM_VB_AINV 0
#define PLM_VB_W 36
#define PLM_VB_FB 42
#define PLM_VB_GF 132
#define PLM_VB_DOUBLES 240

struct NodeArgs {
  const PlmModel* M;
  const PlmLayout* L;
  const PlmNodeType* T;
  const int16_t* lut;             // this node type's lut
  const PlmConstEntry* consts;    // this node type's constant entries
  const double* xs;               // [dx_i | u_i | dx_{i+1}]
  const double* p;                // parameter vector of the instance
  int node;
  double dt;
  int want_jac;
};

struct LaneState {
  double J[6], Vj[6], Vp[6], Ap[6];
  double Jq[6];
  double w[6], y[6], dFv[6], dFq[6], dFqn[6];
  double tau;
  double Tq[6], Tv[6];   // centroidal_vel without base inputs: d v_b / d dq_lane, d v_b / d v_lane
};

#ifndef PLM_EMIT_HINTS
#define PLM_EMIT_HINTS 1
#endif
// Jacobian entry store: written once and never read by this kernel -> streaming store on the device
PLM_HD void store_j(double* dst, double val) {
#if defined(__CUDA_ARCH__) && PLM_EMIT_HINTS
  __stcs(dst, val);
#else
  *dst = val;
#endif
}
PLM_HD void emit(const NodeWs& ws, const NodeArgs& A, int src, int idx, double val) {
#if defined(__CUDA_ARCH__) && PLM_EMIT_HINTS
  // read-only path: the lookups of a run of emits can be issued ahead of the stores of the earlier ones; the table is
  // asked to stay in L1 (evict_last) and the entries, written once and never read here, stream past it (st.global.cs)
  short pos16;
  asm("ld.global.nc.L1::evict_last.s16 %0, [%1];" : "=h"(pos16) : "l"(A.lut + A.T->src_off[src] + idx));
  const int pos = pos16;
  if (pos >= 0) store_j(ws.J + pos, val);
#elif defined(__CUDA_ARCH__)
  int pos = __ldg(A.lut + A.T->src_off[src] + idx);
  if (pos >= 0) ws.J[pos] = val;
#else
  int pos = A.lut[A.T->src_off[src] + idx];
  if (pos >= 0) ws.J[pos] = val;
#endif
}

PLM_HD bool kind_has_state_v(int kind) { return kind != PLM_CENTROIDAL_VEL; }

// Node time step dt_i = dt_min * gamma^i, gamma = (dt_max/dt_min)^(1/(N-1))   (optimization/ocp.py:71-74)
// (gamma^i by binary exponentiation: a second pow() per node evaluation was 4 % of the kernel's instructions; the two
// forms differ by a few ulp, far inside the 1e-9 tolerance of the rows)
PLM_HD double node_dt(const PlmLayout& L, const double* p, int i) {
  double dt_min = p[L.p_dt_min], dt_max = p[L.p_dt_max];
  double gamma = pow(dt_max / dt_min, 1.0 / (double)(L.nodes - 1));
  double gi = 1.0, sq = gamma;
  for (int e = i; e > 0; e >>= 1) {
    if (e & 1) gi *= sq;
    sq *= sq;
  }
  return dt_min * gi;
}

// ---------------------------------------------------------------------------------------------
// Phase A: stage per-column coordinates and joint sin/cos.   lane-parallel, no dependencies.
// ---------------------------------------------------------------------------------------------
template <int KIND>
PLM_HD void node_phase_a(NodeWs& ws, const NodeArgs& A, int lane) {
  const PlmModel& M = *A.M;
  const PlmLayout& L = *A.L;
  const double* x_init = A.p + L.p_x_init;
  const int nv = M.nv;
  const int qoff = (KIND == PLM_CENTROIDAL_VEL) ? 6 : 0;       // x_init = [h | q] for centroidal_vel
  const int dqoff = (KIND == PLM_CENTROIDAL_VEL) ? 6 : 0;      // dx = [dh | dq]
  const double* dx = A.xs;
  const double* u = A.xs + L.ndx;
  if (lane < L.nf) ws.fx[lane] = u[L.f_idx + lane];
  if (lane < nv) {
    ws.cq[lane] = dx[dqoff + lane];
    // without base inputs the leading block of U holds the joint part only; the base part starts at zero and is
    // solved for after the first pass (node_phase_base_solve)
    double ulead = L.nobase ? (lane >= 6 ? u[lane - 6] : 0.0) : ((KIND == PLM_WHOLE_BODY_ABA) ? 0.0 : u[lane]);
    if (KIND == PLM_WHOLE_BODY_RNEA && L.noacc) {
      // no acceleration inputs: a = (v_{i+1} - v_i) / dt_i = (dv_{i+1} - dv_i) / dt_i   (ocp_whole_body_rnea.py:183-191)
      const double* dxn = A.xs + L.ndx + A.T->nu;
      ulead = (dxn[nv + lane] - dx[nv + lane]) / A.dt;
    }
    if (KIND == PLM_CENTROIDAL_VEL) {
      ws.cv[lane] = ulead;
      ws.ca[lane] = 0.0;
    } else {
      ws.cv[lane] = x_init[M.nq + lane] + dx[nv + lane];
      ws.ca[lane] = ulead;
    }
  }
  int b = lane + 1;
  if (b < M.nbody) {
    double ang = x_init[qoff + 7 + lane] + dx[dqoff + 6 + lane];
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(ang, &s, &c);
#else
    s = sin(ang); c = cos(ang);
#endif
    ws.sc[b][0] = s;
    ws.sc[b][1] = c;
  }
  if (lane == 0) {
    double R0[9], E[9];
    quat_to_R(x_init + qoff + 3, R0);
    const double* w = dx + dqoff + 3;
    exp3(w, E);
    matmul3(R0, E, ws.Rb);
    for (int i = 0; i < 9; ++i) ws.Rinit[i] = R0[i];
    jr3(w, ws.jr);
  }
}

// Rotate R (row-major) by a revolute joint: R <- R * Rot(axis, angle)
PLM_HD void apply_joint_rot(double* R, int axtype, const double* axis, double s, double c) {
  if (axtype == 0) {        // about x: columns 1,2
    for (int r = 0; r < 3; ++r) {
      double a = R[3 * r + 1], b = R[3 * r + 2];
      R[3 * r + 1] = c * a + s * b;
      R[3 * r + 2] = -s * a + c * b;
    }
  } else if (axtype == 1) { // about y: columns 2,0
    for (int r = 0; r < 3; ++r) {
      double a = R[3 * r + 2], b = R[3 * r + 0];
      R[3 * r + 2] = c * a + s * b;
      R[3 * r + 0] = -s * a + c * b;
    }
  } else if (axtype == 2) { // about z: columns 0,1
    for (int r = 0; r < 3; ++r) {
      double a = R[3 * r + 0], b = R[3 * r + 1];
      R[3 * r + 0] = c * a + s * b;
      R[3 * r + 1] = -s * a + c * b;
    }
  } else {                  // Rodrigues: Rot = I + s K + (1-c) K^2
    double Rot[9];
    double x = axis[0], y = axis[1], z = axis[2], t = 1.0 - c;
    Rot[0] = c + t * x * x;     Rot[1] = t * x * y - s * z; Rot[2] = t * x * z + s * y;
    Rot[3] = t * x * y + s * z; Rot[4] = c + t * y * y;     Rot[5] = t * y * z - s * x;
    Rot[6] = t * x * z - s * y; Rot[7] = t * y * z + s * x; Rot[8] = c + t * z * z;
    matmul3(R, Rot, R);
  }
}

// q-direction of a column: base angular columns are mixed by the right Jacobian of Exp (dq -> local tangent), base
// linear columns have no effect on any row, joint columns are their own direction.
PLM_HD void lane_q_direction(const NodeWs& ws, LaneState& st, int lane, int body) {
  if (body == 0) {
    st.Jq[0] = st.Jq[1] = st.Jq[2] = 0.0;
    if (lane < 3) { st.Jq[3] = st.Jq[4] = st.Jq[5] = 0.0; }
    else {
      double e[3] = {ws.jr[lane - 3], ws.jr[3 + lane - 3], ws.jr[6 + lane - 3]};   // column (lane-3) of Jr
      matvec3(ws.Rb, e, st.Jq + 3);
    }
  } else {
    for (int i = 0; i < 6; ++i) st.Jq[i] = st.J[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Phase B: every lane walks its own chain (forward kinematics + velocity/acceleration recursion),
// forms the world inertia of its body, the body force and Coriolis block, and the owner lane stores
// the body record.  Contact and arm frames are recorded by the lane that owns their parent body.
// ---------------------------------------------------------------------------------------------
template <int KIND>
PLM_HD void node_phase_b(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  const PlmModel& M = *A.M;
  if (lane >= M.nv) return;
  const double* Rb = ws.Rb;
  double R[9], p[3] = {0, 0, 0};
  for (int i = 0; i < 9; ++i) R[i] = Rb[i];
  // base spatial velocity / acceleration (world axes, about the base origin)
  double Vj[6], Aj[6], Vp[6] = {0, 0, 0, 0, 0, 0}, Ap[6] = {0, 0, M.gravity_z, 0, 0, 0};
  matvec3(Rb, ws.cv, Vj);
  matvec3(Rb, ws.cv + 3, Vj + 3);
  matvec3(Rb, ws.ca, Aj);
  matvec3(Rb, ws.ca + 3, Aj + 3);
  Aj[2] += M.gravity_z;
  double J[6];
  const int body = M.col_body[lane];
  if (body == 0) {
    int k = lane % 3;
    double e[3] = {Rb[k], Rb[3 + k], Rb[6 + k]};
    if (lane < 3) { J[0] = e[0]; J[1] = e[1]; J[2] = e[2]; J[3] = J[4] = J[5] = 0.0; }
    else          { J[0] = J[1] = J[2] = 0.0; J[3] = e[0]; J[4] = e[1]; J[5] = e[2]; }
  }
  const int len = M.chain_len[lane];
  for (int l = 0; l < len; ++l) {
    const int jb = M.chain[lane][l];
    // placement
    double t[3];
    matvec3(R, M.place_p[jb], t);
    p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
    if (M.has_rot[jb]) matmul3(R, M.place_R[jb], R);
    apply_joint_rot(R, M.axtype[jb], M.axis[jb], ws.sc[jb][0], ws.sc[jb][1]);
    double wax[3];
    matvec3(R, M.axis[jb], wax);
    cross3(p, wax, J);
    J[3] = wax[0]; J[4] = wax[1]; J[5] = wax[2];
    const double vq = ws.cv[jb + 5], aq = ws.ca[jb + 5];
    for (int i = 0; i < 6; ++i) { Vp[i] = Vj[i]; Ap[i] = Aj[i]; }
    double bias[6];
    mxm(Vp, J, bias);
    for (int i = 0; i < 6; ++i) {
      Aj[i] += J[i] * aq + bias[i] * vq;
      Vj[i] += J[i] * vq;
    }
  }
  for (int i = 0; i < 6; ++i) { st.J[i] = J[i]; st.Vj[i] = Vj[i]; st.Vp[i] = Vp[i]; st.Ap[i] = Ap[i]; }
  lane_q_direction(ws, st, lane, body);

  const bool owner = (body > 0) || (lane == 0);
  if (!owner) return;
  // world inertia about the base origin
  const double m = M.mass[body];
  double cw[3];
  matvec3(R, M.com[body], cw);
  cw[0] += p[0]; cw[1] += p[1]; cw[2] += p[2];
  double mc[3] = {m * cw[0], m * cw[1], m * cw[2]};
  // Ib = R Ic R^T - m [cw]x^2 = R Ic R^T + m (|cw|^2 I - cw cw^T)
  const double* Ic = M.Ic[body];
  double Icm[9] = {Ic[0], Ic[1], Ic[2], Ic[1], Ic[3], Ic[4], Ic[2], Ic[4], Ic[5]};
  double T1[9], T2[9];
  matmul3(R, Icm, T1);
  // T2 = T1 * R^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T2[3 * i + j] = T1[3 * i] * R[3 * j] + T1[3 * i + 1] * R[3 * j + 1] + T1[3 * i + 2] * R[3 * j + 2];
  double c2 = dot3(cw, cw);
  double Ib[6];
  Ib[0] = T2[0] + m * (c2 - cw[0] * cw[0]);
  Ib[1] = 0.5 * (T2[1] + T2[3]) - m * cw[0] * cw[1];
  Ib[2] = 0.5 * (T2[2] + T2[6]) - m * cw[0] * cw[2];
  Ib[3] = T2[4] + m * (c2 - cw[1] * cw[1]);
  Ib[4] = 0.5 * (T2[5] + T2[7]) - m * cw[1] * cw[2];
  Ib[5] = T2[8] + m * (c2 - cw[2] * cw[2]);
  double H[6], F[6], t6[6];
  inertia_mul(m, mc, Ib, Vj, H);
  inertia_mul(m, mc, Ib, Aj, F);
  mxf(Vj, H, t6);
  for (int i = 0; i < 6; ++i) F[i] += t6[i];
  // external forces on this body (world force f_k at world point p_k)
  for (int k = 0; k < M.ncontact; ++k) {
    if (M.contact_body[k] != body) continue;
    double pk[3];
    matvec3(R, M.contact_off[k], pk);
    pk[0] += p[0]; pk[1] += p[1]; pk[2] += p[2];
    const double* fk = ws.fx + 3 * k;
    double n[3];
    cross3(pk, fk, n);
    F[0] -= fk[0]; F[1] -= fk[1]; F[2] -= fk[2];
    F[3] -= n[0];  F[4] -= n[1];  F[5] -= n[2];
    for (int i = 0; i < 3; ++i) ws.con[k][i] = pk[i];
    for (int i = 0; i < 6; ++i) ws.con[k][3 + i] = Vj[i];
  }
  if (M.arm_body == body) {
    double pk[3];
    matvec3(R, M.arm_off, pk);
    for (int i = 0; i < 3; ++i) ws.arm[i] = pk[i] + p[i];
    for (int i = 0; i < 6; ++i) ws.arm[3 + i] = Vj[i];
  }
  // Coriolis block B22 = [w]x Ib - Ib [w]x - [v]x [mc]x - [mc]x [v]x - [H_ang]x
  const double* v = Vj;
  const double* w = Vj + 3;
  double Ibm[9] = {Ib[0], Ib[1], Ib[2], Ib[1], Ib[3], Ib[4], Ib[2], Ib[4], Ib[5]};
  double B[9];
  // ([w]x Ib)_{ij} = sum_k eps_{i a k} w_a Ib_{kj} : row i = w x (column j of Ib)  -> compute per column
  for (int j = 0; j < 3; ++j) {
    double colj[3] = {Ibm[j], Ibm[3 + j], Ibm[6 + j]};
    double wc[3];
    cross3(w, colj, wc);                // column j of [w]x Ib
    B[0 + j] = wc[0]; B[3 + j] = wc[1]; B[6 + j] = wc[2];
  }
  // - Ib [w]x : row i of (Ib [w]x) = (Ib row i) x ... (r [w]x) = r x w  => subtract (row_i x w)
  for (int i = 0; i < 3; ++i) {
    double rowi[3] = {Ibm[3 * i], Ibm[3 * i + 1], Ibm[3 * i + 2]};
    double rw[3];
    cross3(rowi, w, rw);
    B[3 * i] -= rw[0]; B[3 * i + 1] -= rw[1]; B[3 * i + 2] -= rw[2];
  }
  // [a]x [b]x = b a^T - (a.b) I   =>  [v]x[mc]x + [mc]x[v]x = mc v^T + v mc^T - 2 (v.mc) I
  double vm = dot3(v, mc);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[3 * i + j] -= mc[i] * v[j] + v[i] * mc[j];
  B[0] += 2.0 * vm; B[4] += 2.0 * vm; B[8] += 2.0 * vm;
  // - [H_ang]x
  B[1] += H[5];  B[2] -= H[4];
  B[3] -= H[5];  B[5] += H[3];
  B[6] += H[4];  B[7] -= H[3];
  double* r = ws.rec[body];
  for (int i = 0; i < 6; ++i) r[PLM_REC_F + i] = F[i];
  r[PLM_REC_M] = m;
  for (int i = 0; i < 3; ++i) r[PLM_REC_MC + i] = mc[i];
  for (int i = 0; i < 6; ++i) r[PLM_REC_IB + i] = Ib[i];
  for (int i = 0; i < 6; ++i) r[PLM_REC_H + i] = H[i];
  for (int i = 0; i < 9; ++i) r[PLM_REC_B + i] = B[i];
}

// ---------------------------------------------------------------------------------------------
// Phase C: subtree composites, leaves first; lanes run over the record entries.
// ---------------------------------------------------------------------------------------------
// Lane l only ever touches entry l of the records, so the steps of one lane are ordered by program order alone: the
// whole sweep is one phase without warp barriers between the steps.
PLM_HD void node_phase_c(NodeWs& ws, const PlmModel& M, int lane) {
  if (lane >= PLM_REC) return;
  for (int step = 0; step < M.nbody - 1; ++step) {
    const int child = M.body_order[step];
    const int par = M.parent[child];
    ws.rec[par][lane] += ws.rec[child][lane];
  }
}

PLM_HD void bc_mul(const double* rec, const double* mot, double* o) {   // B^C * mot
  const double* P = rec + PLM_REC_H;   // composite linear momentum
  const double* B = rec + PLM_REC_B;
  const double* al = mot + 3;
  double t[3];
  cross3(P, al, t);
  o[0] = -2.0 * t[0]; o[1] = -2.0 * t[1]; o[2] = -2.0 * t[2];
  matvec3(B, al, o + 3);
}
PLM_HD void bct_mul(const double* rec, const double* mot, double* o) {  // (B^C)^T * mot
  const double* P = rec + PLM_REC_H;
  const double* B = rec + PLM_REC_B;
  double t[3];
  cross3(P, mot, t);
  o[0] = o[1] = o[2] = 0.0;
  matTvec3(B, mot + 3, o + 3);
  o[3] += 2.0 * t[0]; o[4] += 2.0 * t[1]; o[5] += 2.0 * t[2];
}

// wrench of a force g applied at point pk
PLM_HD void point_wrench(const double* pk, const double* g, double* o) {
  o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
  cross3(pk, g, o + 3);
}

// ---------------------------------------------------------------------------------------------
// Phase D: per-column derivative directions from the composites of the column's own body.
// ---------------------------------------------------------------------------------------------
template <int KIND>
PLM_HD void node_phase_d(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  const PlmModel& M = *A.M;
  if (lane >= M.nv) return;
  const int body = M.col_body[lane];
  const double* rec = ws.rec[body];
  const double* FC = rec + PLM_REC_F;
  const double mC = rec[PLM_REC_M];
  const double* mcC = rec + PLM_REC_MC;
  const double* IbC = rec + PLM_REC_IB;
  st.tau = dot6(st.J, FC);
  if (!A.want_jac) return;
  double Jq[6];
  for (int i = 0; i < 6; ++i) Jq[i] = st.Jq[i];
  double phi[6], chi[6], psi[6], t6[6], VV[6];
  mxm(st.Vp, Jq, phi);
  mxm(st.Ap, Jq, chi);
  mxm(st.Vp, phi, t6);
  for (int i = 0; i < 6; ++i) { chi[i] += t6[i]; VV[i] = st.Vj[i] + st.Vp[i]; }
  mxm(VV, st.J, psi);
  inertia_mul(mC, mcC, IbC, st.J, st.w);
  bct_mul(rec, st.J, st.y);
  inertia_mul(mC, mcC, IbC, psi, st.dFv);
  bc_mul(rec, st.J, t6);
  for (int i = 0; i < 6; ++i) st.dFv[i] += t6[i];
  inertia_mul(mC, mcC, IbC, chi, st.dFqn);
  bc_mul(rec, phi, t6);
  for (int i = 0; i < 6; ++i) st.dFqn[i] += t6[i];
  const unsigned mask = M.col_contacts[lane];
  for (int k = 0; k < M.ncontact; ++k) {
    if (!((mask >> k) & 1u)) continue;
    double g[3], wr[6];
    cross3(Jq + 3, ws.fx + 3 * k, g);
    point_wrench(ws.con[k], g, wr);
    for (int i = 0; i < 6; ++i) st.dFqn[i] += wr[i];
  }
  mxf(Jq, FC, st.dFq);
  for (int i = 0; i < 6; ++i) st.dFq[i] += st.dFqn[i];
  double* c = ws.col[lane];
  for (int i = 0; i < 6; ++i) { c[i] = st.J[i]; c[6 + i] = phi[i]; c[12 + i] = chi[i]; c[18 + i] = psi[i]; }
  if (lane < 6) { ws.jqb[lane][0] = Jq[3]; ws.jqb[lane][1] = Jq[4]; ws.jqb[lane][2] = Jq[5]; }
}

PLM_HD void shift_to(const double* F, const double* c, double* o) {   // wrench about the origin -> about c
  double t[3];
  cross3(c, F, t);
  o[0] = F[0]; o[1] = F[1]; o[2] = F[2];
  o[3] = F[3] - t[0]; o[4] = F[4] - t[1]; o[5] = F[5] - t[2];
}

// ---------------------------------------------------------------------------------------------
// Formulations without base inputs (include_base = False).  The base part w of the velocity (centroidal_vel) or of the
// acceleration (centroidal_acc, whole_body_acc) follows from the six dynamics-gap rows, which are affine in it:
//   gaps(w) = A_b w + gaps(0),   w = -A_b^-1 gaps(0),   d w / d theta = -A_b^-1 d gaps / d theta
// (dynamics_centroidal_vel.py:73-89 base_vel; dynamics_centroidal_acc.py:43-82, dynamics_whole_body_acc.py:43-83 base_acc).
// Pass 1 (phases A-C with w = 0) gives gaps(0) and A_b = d gaps / d w (lanes 0..5 own one column each); pass 2 repeats
// phases B, C with w in place.  The gap rows themselves are not part of these formulations.
// ---------------------------------------------------------------------------------------------
template <int KIND>
PLM_HD void node_phase_base_cols(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  if (lane >= 6) return;
  const PlmLayout& L = *A.L;
  const double* root = ws.rec[0];
  const double mC = root[PLM_REC_M];
  const double* mcC = root + PLM_REC_MC;
  double w[6];
  inertia_mul(mC, mcC, root + PLM_REC_IB, st.J, w);      // I^C_0 J_lane
  double* Ab = ws.vb + PLM_VB_AINV;
  double* g0 = ws.vb + PLM_VB_W;
  if (KIND == PLM_WHOLE_BODY_ACC) {
    // M_bb[c][lane] = J_c . I^C_0 J_lane ;  gaps(0)[lane] = J_lane . F^C_0
    for (int c = 0; c < 6; ++c) {
      const int k = c % 3;
      const double e[3] = {ws.Rb[k], ws.Rb[3 + k], ws.Rb[6 + k]};
      Ab[6 * c + lane] = (c < 3) ? dot3(e, w) : dot3(e, w + 3);
    }
    g0[lane] = dot6(st.J, root + PLM_REC_F);
  } else {
    const double com[3] = {mcC[0] / mC, mcC[1] / mC, mcC[2] / mC};
    double o[6];
    shift_to(w, com, o);
    for (int r = 0; r < 6; ++r) Ab[6 * r + lane] = o[r];
    if (lane == 0) {
      shift_to(KIND == PLM_CENTROIDAL_ACC ? root + PLM_REC_F : root + PLM_REC_H, com, o);
      if (KIND == PLM_CENTROIDAL_VEL) {
        const double* x_init = A.p + L.p_x_init;
        for (int r = 0; r < 6; ++r) o[r] -= mC * (x_init[r] + A.xs[r]);
      }
      for (int r = 0; r < 6; ++r) g0[r] = o[r];
    }
  }
}

// Lane 0: A_b^-1 by Gauss-Jordan elimination with partial pivoting (A_b of the centroidal forms is not symmetric: rows
// in world axes about the centre of mass, columns in base axes), then w = -A_b^-1 gaps(0) into the coordinate arrays.
template <int KIND>
PLM_HD void node_phase_base_solve(NodeWs& ws, int lane) {
  if (lane != 0) return;
  double* Ab = ws.vb + PLM_VB_AINV;
  double* g0 = ws.vb + PLM_VB_W;
  double* aug = ws.vb + PLM_VB_GF;      // [6][12] scratch (the force columns are filled later)
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) { aug[12 * r + c] = Ab[6 * r + c]; aug[12 * r + 6 + c] = (r == c) ? 1.0 : 0.0; }
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(aug[12 * k + k]);
    for (int r = k + 1; r < 6; ++r) {
      const double v = fabs(aug[12 * r + k]);
      if (v > best) { best = v; piv = r; }
    }
    if (piv != k)
      for (int c = 0; c < 12; ++c) { const double t = aug[12 * k + c]; aug[12 * k + c] = aug[12 * piv + c]; aug[12 * piv + c] = t; }
    const double inv = 1.0 / aug[12 * k + k];
    for (int c = 0; c < 12; ++c) aug[12 * k + c] *= inv;
    for (int r = 0; r < 6; ++r) {
      if (r == k) continue;
      const double f = aug[12 * r + k];
      for (int c = 0; c < 12; ++c) aug[12 * r + c] -= f * aug[12 * k + c];
    }
  }
  double* dst = (KIND == PLM_CENTROIDAL_VEL) ? ws.cv : ws.ca;
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
    for (int c = 0; c < 6; ++c) { Ab[6 * r + c] = aug[12 * r + 6 + c]; acc += aug[12 * r + 6 + c] * g0[c]; }
    dst[r] = -acc;
  }
}

// Base-integrator rows of a formulation without base inputs: row c (c < 6) gets dt (A_b^-1 o)[c] in the column owned by
// this lane, where o = d gaps / d (that column); `diag`: the column is the row's own state entry (-1 folded in).
PLM_HD void emit_base_rows(const NodeWs& ws, const NodeArgs& A, int src, int stride, int col, const double* o, bool diag) {
  const double* Ainv = ws.vb + PLM_VB_AINV;
  for (int c = 0; c < 6; ++c) {
    double t = A.dt * dot6(Ainv + 6 * c, o);
    if (diag && c == col) t -= 1.0;
    emit(ws, A, src, c * stride + col, t);
  }
}

// ---------------------------------------------------------------------------------------------
// Phase E: dynamics rows (values and Jacobian entries).
// ---------------------------------------------------------------------------------------------
template <int KIND, bool NOBASE>
PLM_HD void node_phase_e(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  const PlmModel& M = *A.M;
  const PlmLayout& L = *A.L;
  const PlmNodeType& T = *A.T;
  const int nv = M.nv, nf = L.nf;
  if (lane >= nv) return;
  const double* u = A.xs + L.ndx;
  const int body = M.col_body[lane];
  const unsigned mask = M.col_contacts[lane];

  if (KIND == PLM_WHOLE_BODY_RNEA || KIND == PLM_WHOLE_BODY_ACC) {
    // ---- values: tau_rnea rows
    if (!NOBASE) {
      if (lane < 6) ws.g[T.row_dyn + lane] = st.tau;
      else if (T.joint_rows) ws.g[T.row_tauj + lane - 6] = st.tau - u[L.tau_idx + lane - 6];
    }
    if (!A.want_jac) return;
    const bool jr = T.joint_rows != 0;
    // ancestors-or-self columns: base 0..5, then chain
    const int len = M.chain_len[lane];
    const int nanc = 6 + (jr ? len : 0);
    double oa[6], ov[6], oq[6];      // without base inputs: d(base rows)/d(this lane's columns)
    for (int a = 0; a < nanc; ++a) {
      const int c = (a < 6) ? a : (M.chain[lane][a - 6] + 5);
      const double* cr = ws.col[c];
      const bool same = (M.col_body[c] == body);
      if (same && c != lane && body != 0) continue;   // (cannot happen for 1-dof joints)
      double va = dot6(cr, st.w);
      double vv = dot6(cr, st.dFv);
      double vq = dot6(cr, same ? st.dFqn : st.dFq);
      if (NOBASE) { oa[a] = va; ov[a] = vv; oq[a] = vq; continue; }
      if (KIND == PLM_WHOLE_BODY_RNEA && L.noacc) {      // a = (dv_next - dv) / dt: the a-column moves to dv (-1/dt) and dv_next (+1/dt)
        const double vn = va / A.dt;
        emit(ws, A, PLM_SRC_TN, c * nv + lane, vn);
        emit(ws, A, PLM_SRC_TV, c * nv + lane, vv - vn);
      } else {
        emit(ws, A, PLM_SRC_TA, c * nv + lane, va);
        emit(ws, A, PLM_SRC_TV, c * nv + lane, vv);
      }
      emit(ws, A, PLM_SRC_TQ, c * nv + lane, vq);
      if (!same && jr) {
        // row = this lane's column (descendant), column = ancestor c
        double tv = dot6(st.w, cr + 18) + dot6(st.y, cr);
        double tq = dot6(st.w, cr + 12) + dot6(st.y, cr + 6);
        for (int k = 0; k < M.ncontact; ++k) {
          if (!((mask >> k) & 1u)) continue;
          double g[3], wr[6];
          cross3(c < 6 ? ws.jqb[c] : cr + 3, ws.fx + 3 * k, g);
          point_wrench(ws.con[k], g, wr);
          tq += dot6(st.J, wr);
        }
        if (KIND == PLM_WHOLE_BODY_RNEA && L.noacc) {
          const double vn = va / A.dt;
          emit(ws, A, PLM_SRC_TN, lane * nv + c, vn);
          emit(ws, A, PLM_SRC_TV, lane * nv + c, tv - vn);
        } else {
          emit(ws, A, PLM_SRC_TA, lane * nv + c, va);
          emit(ws, A, PLM_SRC_TV, lane * nv + c, tv);
        }
        emit(ws, A, PLM_SRC_TQ, lane * nv + c, tq);
      }
    }
    if (NOBASE) {
      // base dv-integrator rows dv_next - dv - dt a_b(dq, dv, a_j, f)
      emit_base_rows(ws, A, PLM_SRC_TA, nv, lane, oa, false);
      emit_base_rows(ws, A, PLM_SRC_TV, nv, lane, ov, true);
      emit_base_rows(ws, A, PLM_SRC_TQ, nv, lane, oq, false);
    }
    // forces: d tau_lane / d f_k = -(J_lin + J_ang x p_k)
    if (lane < 6 || jr) {
      for (int k = 0; k < M.ncontact; ++k) {
        if (!((mask >> k) & 1u)) continue;
        double jk[3];
        cross3(st.J + 3, ws.con[k], jk);
        for (int t = 0; t < 3; ++t) {
          if (NOBASE) ws.vb[PLM_VB_GF + 6 * (3 * k + t) + lane] = -(st.J[t] + jk[t]);      // row lane of force column (k, t)
          else emit(ws, A, PLM_SRC_TF, lane * nf + 3 * k + t, -(st.J[t] + jk[t]));
        }
      }
    }
  }

  if (KIND == PLM_CENTROIDAL_ACC || KIND == PLM_CENTROIDAL_VEL) {
    const double* root = ws.rec[0];
    const double Mtot = root[PLM_REC_M];
    double com[3] = {root[PLM_REC_MC] / Mtot, root[PLM_REC_MC + 1] / Mtot, root[PLM_REC_MC + 2] / Mtot};
    const double* tot = (KIND == PLM_CENTROIDAL_ACC) ? (root + PLM_REC_F) : (root + PLM_REC_H);
    const double* x_init = A.p + L.p_x_init;
    if (lane == 0) {
      double gsh[6];
      shift_to(tot, com, gsh);
      if (KIND == PLM_CENTROIDAL_VEL) {
        // gaps = A v - m h,  h = h_init + dh
        for (int r = 0; r < 6; ++r) gsh[r] -= Mtot * (x_init[r] + A.xs[r]);
      }
      if (!NOBASE) for (int r = 0; r < 6; ++r) ws.g[T.row_dyn + r] = gsh[r];
    }
    if (KIND == PLM_CENTROIDAL_VEL) {
      // h_dot = [sum f + m g; sum (p_k - c) x f_k] / m ; rows dh_next - (dh + h_dot dt)
      if (lane == 1) {
        double hd[6] = {0, 0, -M.gravity_z * Mtot, 0, 0, 0};
        for (int k = 0; k < M.ncontact; ++k) {
          const double* fk = ws.fx + 3 * k;
          double r3[3] = {ws.con[k][0] - com[0], ws.con[k][1] - com[1], ws.con[k][2] - com[2]};
          hd[0] += fk[0]; hd[1] += fk[1]; hd[2] += fk[2];
          cross3_acc(r3, fk, hd + 3);
        }
        const double* dxn = A.xs + L.ndx + T.nu;
        for (int r = 0; r < 6; ++r) ws.g[T.row_int + r] = dxn[r] - (A.xs[r] + hd[r] / Mtot * A.dt);
      }
    }
    if (!A.want_jac) return;
    const double* rec = ws.rec[body];
    const double mC = rec[PLM_REC_M];
    const double* mcC = rec + PLM_REC_MC;
    // d com / d q along Jq
    double dc[3], t3[3];
    cross3(st.Jq + 3, mcC, t3);
    for (int i = 0; i < 3; ++i) dc[i] = (mC * st.Jq[i] + t3[i]) / Mtot;
    double o[6];
    if (KIND == PLM_CENTROIDAL_ACC) {
      shift_to(st.dFq, com, o);
      cross3(dc, tot, t3);
      for (int r = 0; r < 3; ++r) o[3 + r] -= t3[r];
      if (NOBASE) emit_base_rows(ws, A, PLM_SRC_TQ, nv, lane, o, false);
      else for (int r = 0; r < 6; ++r) emit(ws, A, PLM_SRC_TQ, r * nv + lane, o[r]);
      shift_to(st.dFv, com, o);
      if (NOBASE) emit_base_rows(ws, A, PLM_SRC_TV, nv, lane, o, true);
      else for (int r = 0; r < 6; ++r) emit(ws, A, PLM_SRC_TV, r * nv + lane, o[r]);
      shift_to(st.w, com, o);
      if (NOBASE) emit_base_rows(ws, A, PLM_SRC_TA, nv, lane, o, false);
      else for (int r = 0; r < 6; ++r) emit(ws, A, PLM_SRC_TA, r * nv + lane, o[r]);
    } else {
      // d(shift_c H)/dq = shift(Jq x* H^C + I^C phi_q) - (0, dc x H_lin);  d/dv = shift(I^C J)
      double dH[6], t6[6];
      mxf(st.Jq, rec + PLM_REC_H, dH);
      inertia_mul(mC, mcC, rec + PLM_REC_IB, ws.col[lane] + 6, t6);
      for (int i = 0; i < 6; ++i) dH[i] += t6[i];
      shift_to(dH, com, o);
      cross3(dc, tot, t3);
      for (int r = 0; r < 3; ++r) o[3 + r] -= t3[r];
      if (NOBASE) {
        // base dq-integrator rows dq_next - dq - dt v_b(dh, dq, v_j); the directions d v_b / d(column) are kept for the
        // foot / arm velocity rows (node_phase_fjac)
        const double* Ainv = ws.vb + PLM_VB_AINV;
        emit_base_rows(ws, A, PLM_SRC_TQ, nv, lane, o, true);
        for (int c = 0; c < 6; ++c) st.Tq[c] = -dot6(Ainv + 6 * c, o);
        shift_to(st.w, com, o);
        emit_base_rows(ws, A, PLM_SRC_TV, nv, lane, o, false);
        for (int c = 0; c < 6; ++c) st.Tv[c] = -dot6(Ainv + 6 * c, o);
        if (lane < 6)      // d gaps / d dh_lane = -m e_lane
          for (int c = 0; c < 6; ++c) emit(ws, A, PLM_SRC_IH, c * 6 + lane, -A.dt * Mtot * Ainv[6 * c + lane]);
      } else {
        for (int r = 0; r < 6; ++r) emit(ws, A, PLM_SRC_TQ, r * nv + lane, o[r]);
        shift_to(st.w, com, o);
        for (int r = 0; r < 6; ++r) emit(ws, A, PLM_SRC_TV, r * nv + lane, o[r]);
      }
      // d(h_dot)/dq: angular rows sum_k (dp_k - dc) x f_k / m, times -dt for the integrator row
      double acc[3] = {0, 0, 0};
      for (int k = 0; k < M.ncontact; ++k) {
        double dp[3] = {-dc[0], -dc[1], -dc[2]};
        if ((mask >> k) & 1u) {
          cross3(st.Jq + 3, ws.con[k], t3);
          for (int i = 0; i < 3; ++i) dp[i] += st.Jq[i] + t3[i];
        }
        cross3_acc(dp, ws.fx + 3 * k, acc);
      }
      for (int r = 0; r < 3; ++r) emit(ws, A, PLM_SRC_XQ, (3 + r) * nv + lane, -A.dt * acc[r] / Mtot);
    }
    // force columns: one contact per lane
    if (lane < M.ncontact) {
      const int k = lane;
      double r3[3] = {ws.con[k][0] - com[0], ws.con[k][1] - com[1], ws.con[k][2] - com[2]};
      for (int t = 0; t < 3; ++t) {
        double e[3] = {0, 0, 0};
        e[t] = 1.0;
        double n[3];
        cross3(r3, e, n);
        if (KIND == PLM_CENTROIDAL_ACC && NOBASE) {
          const double of[6] = {t == 0 ? -1.0 : 0.0, t == 1 ? -1.0 : 0.0, t == 2 ? -1.0 : 0.0, -n[0], -n[1], -n[2]};
          emit_base_rows(ws, A, PLM_SRC_TF, nf, 3 * k + t, of, false);
        } else if (KIND == PLM_CENTROIDAL_ACC) {
          emit(ws, A, PLM_SRC_TF, t * nf + 3 * k + t, -1.0);
          for (int r = 0; r < 3; ++r) emit(ws, A, PLM_SRC_TF, (3 + r) * nf + 3 * k + t, -n[r]);
        } else {
          emit(ws, A, PLM_SRC_XF, t * nf + 3 * k + t, -A.dt / Mtot);
          for (int r = 0; r < 3; ++r) emit(ws, A, PLM_SRC_XF, (3 + r) * nf + 3 * k + t, -A.dt * n[r] / Mtot);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Phase F: foot-velocity and arm-velocity rows (lane = column), shared rows and constants.
// ---------------------------------------------------------------------------------------------
// spline z-velocity of utils/gait_sequence.py:96-133
PLM_HD double spline_vel_z(double phase, double T, double h, double v_lo, double v_td) {
  double mid = T / 2;
  double t = phase * T;
  double t0, t1, p0, v0, p1, v1;
  if (phase < 0.5) { t0 = 0; t1 = mid; p0 = 0; v0 = v_lo; p1 = h; v1 = 0; }
  else             { t0 = mid; t1 = T; p0 = h; v0 = 0; p1 = 0; v1 = v_td; }
  double dt = t1 - t0, dpos = p1 - p0, dvel = v1 - v0;
  double c1 = v0 * dt;
  double c2 = -(3.0 * v0 + dvel) * dt + 3.0 * dpos;
  double c3 = (2.0 * v0 + dvel) * dt - 2.0 * dpos;
  double tn = (t - t0) / dt;
  return (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dt;
}

// Foot-velocity and arm-velocity Jacobian entries, lane = column.
// MODE 0: the rows depend on the lane's own columns only (kinematic-tree pattern).
// MODE 1 / 2 (centroidal_vel without base inputs): the frame velocities also depend on every column through
// v_b(dh, dq, v_j).  MODE 1 (lanes 0..5, before phase E): record d(row)/d v_b[lane]; MODE 2 (after phase E): emit
// d(row)/d(column) = direct part + sum_j d(row)/d v_b[j] * d v_b[j]/d(column) for every row and column.
template <int KIND, int MODE>
PLM_HD void node_phase_fjac(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  const PlmModel& M = *A.M;
  const PlmLayout& L = *A.L;
  const PlmNodeType& T = *A.T;
  const int nv = M.nv;
  if (!A.want_jac || lane >= nv || !T.state_rows) return;
  const double* contact = A.p + L.p_contact + 4 * A.node;
  double* fb = (MODE != 0) ? ws.vb + PLM_VB_FB : nullptr;
  const double* Ainv = (MODE == 2) ? ws.vb + PLM_VB_AINV : nullptr;
  const double Mtot = (MODE == 2) ? ws.rec[0][PLM_REC_M] : 0.0;
  if (MODE == 1 && lane >= 6) return;
  // MODE 2: `row` indexes the source block (foot rows 3 k + r, arm rows r), `frow` the recorded coefficients
  auto chain = [&](int frow, int row, int srcq, int srcv, int srch, double dq_direct, double dv_direct) {
    const double* f6 = fb + 6 * frow;
    emit(ws, A, srcq, row * nv + lane, dq_direct + dot6(f6, st.Tq));
    emit(ws, A, srcv, row * nv + lane, dv_direct + dot6(f6, st.Tv));
    if (lane < 6) {      // column dh_lane: d v_b / d dh_lane = m A_b^-1[:, lane]
      double acc = 0.0;
      for (int j = 0; j < 6; ++j) acc += f6[j] * Ainv[6 * j + lane];
      emit(ws, A, srch, row * 6 + lane, Mtot * acc);
    }
  };
  const unsigned mask = M.col_contacts[lane];
  for (int k = 0; k < M.nfeet; ++k) {
    const bool in = ((mask >> k) & 1u) != 0;
    if (MODE != 2 && !in) continue;
    const double* ck = ws.con[k];
    const double c = contact[k];
    const double sc3[3] = {c, c, 1.0};
    double jk[3] = {0, 0, 0}, o[3] = {0, 0, 0};
    if (in) {
      double jq[3], D[6], x6[6], t3[3];
      cross3(st.J + 3, ck, jk);
      cross3(st.Jq + 3, ck, jq);
      for (int i = 0; i < 3; ++i) { jk[i] += st.J[i]; jq[i] += st.Jq[i]; }
      for (int i = 0; i < 6; ++i) D[i] = ck[3 + i] - st.Vp[i];
      mxm(st.Jq, D, x6);
      cross3(x6 + 3, ck, o);
      cross3(ck + 6, jq, t3);
      for (int i = 0; i < 3; ++i) o[i] += x6[i] + t3[i];
    }
    for (int r = 0; r < 3; ++r) {
      if (MODE == 0) {
        emit(ws, A, PLM_SRC_FV, (k * 3 + r) * nv + lane, sc3[r] * jk[r]);
        emit(ws, A, PLM_SRC_FQ, (k * 3 + r) * nv + lane, sc3[r] * o[r]);
      } else if (MODE == 1) fb[6 * (k * 3 + r) + lane] = sc3[r] * jk[r];
      else chain(k * 3 + r, k * 3 + r, PLM_SRC_FQ, PLM_SRC_FV, PLM_SRC_FH, sc3[r] * o[r], sc3[r] * jk[r]);
    }
  }
  if (T.row_arm >= 0) {
    const bool in = ((M.col_arm >> lane) & 1u) != 0;
    if (MODE != 2 && !in) return;
    const double* ak = ws.arm;
    double jk[3] = {0, 0, 0}, o[3] = {0, 0, 0}, jq[3] = {0, 0, 0}, t3[3] = {0, 0, 0};
    if (in) {
      double D[6], x6[6];
      cross3(st.J + 3, ak, jk);
      cross3(st.Jq + 3, ak, jq);
      for (int i = 0; i < 3; ++i) { jk[i] += st.J[i]; jq[i] += st.Jq[i]; }
      for (int i = 0; i < 6; ++i) D[i] = ak[3 + i] - st.Vp[i];
      mxm(st.Jq, D, x6);
      cross3(x6 + 3, ak, o);
      for (int i = 0; i < 3; ++i) o[i] += x6[i];
      cross3(ak + 6, jq, t3);
    }
    // row 2: world z of the full frame velocity
    if (MODE == 0) {
      emit(ws, A, PLM_SRC_AV, 2 * nv + lane, jk[2]);
      emit(ws, A, PLM_SRC_AQ, 2 * nv + lane, o[2] + t3[2]);
    } else if (MODE == 1) {
      fb[6 * (3 * M.nfeet + 0) + lane] = 0.0;      // rows 0, 1: velocity relative to the base, independent of v_b
      fb[6 * (3 * M.nfeet + 1) + lane] = 0.0;
      fb[6 * (3 * M.nfeet + 2) + lane] = jk[2];
    } else chain(3 * M.nfeet + 2, 2, PLM_SRC_AQ, PLM_SRC_AV, PLM_SRC_AH, o[2] + t3[2], jk[2]);
    if (MODE != 1 && in && M.col_body[lane] != 0) {
      // rows 0,1: velocity relative to the base, in base axes (only arm-chain columns contribute)
      double Wb[3], wrel[3], orel[3], ob[3], jb[3];
      matvec3(ws.Rb, ws.cv + 3, Wb);
      for (int i = 0; i < 3; ++i) wrel[i] = ak[6 + i] - Wb[i];
      cross3(wrel, jq, t3);
      for (int i = 0; i < 3; ++i) orel[i] = o[i] + t3[i];
      matTvec3(ws.Rb, orel, ob);
      matTvec3(ws.Rb, jk, jb);
      for (int r = 0; r < 2; ++r) {
        emit(ws, A, PLM_SRC_AV, r * nv + lane, jb[r]);
        emit(ws, A, PLM_SRC_AQ, r * nv + lane, ob[r]);
      }
    }
  }
}

template <int KIND, bool NOBASE>
PLM_HD void node_phase_f(NodeWs& ws, const NodeArgs& A, LaneState& st, int lane) {
  const PlmModel& M = *A.M;
  const PlmLayout& L = *A.L;
  const PlmNodeType& T = *A.T;
  const int nv = M.nv, nf = L.nf, nj = M.nj;
  const double* p = A.p;
  const double* dx = A.xs;
  const double* u = A.xs + L.ndx;
  const double* dxn = A.xs + L.ndx + T.nu;
  const double* x_init = p + L.p_x_init;
  const double* contact = p + L.p_contact + 4 * A.node;   // (4, N) column-major
  const double* swing = p + L.p_swing + 4 * A.node;
  const double dt = A.dt;

  // ---- integrator rows
  if (lane < nv) {
    if (KIND == PLM_CENTROIDAL_VEL) {
      ws.g[T.row_int + 6 + lane] = dxn[6 + lane] - (dx[6 + lane] + ws.cv[lane] * dt);
    } else {
      ws.g[T.row_int + lane] = dxn[lane] - (dx[lane] + ws.cv[lane] * dt);
      if (!L.noacc) ws.g[T.row_int + nv + lane] = dxn[nv + lane] - (dx[nv + lane] + ws.ca[lane] * dt);
    }
  }
  // ---- joint rows: torque bounds, joint position / velocity bounds
  if (lane < nj) {
    if (T.row_taub >= 0) ws.g[T.row_taub + lane] = u[L.tau_idx + lane];
    if (T.row_qj >= 0) {
      const int qoff = (KIND == PLM_CENTROIDAL_VEL) ? 6 : 0;
      ws.g[T.row_qj + lane] = x_init[qoff + 7 + lane] + ws.cq[6 + lane];
      ws.g[T.row_vj + lane] = ws.cv[6 + lane];
    }
  }
  // ---- per-foot force rows (lanes 0..3), external-force rows (lane 4)
  if (lane < M.nfeet) {
    const int k = lane;
    const double c = contact[k];
    const double* f = ws.fx + 3 * k;
    const int r0 = T.row_foot[k];
    const double mu2 = L.mu * L.mu;
    ws.g[r0] = c * f[2];
    ws.g[r0 + 1] = c * (f[0] * f[0] + f[1] * f[1]) - c * mu2 * f[2] * f[2];
    ws.g[r0 + 2] = (1.0 - c) * f[0];
    ws.g[r0 + 3] = (1.0 - c) * f[1];
    ws.g[r0 + 4] = (1.0 - c) * f[2];
    if (T.state_rows) {
      const double* ck = ws.con[k];
      double vf[3];
      cross3(ck + 6, ck, vf);
      vf[0] += ck[3]; vf[1] += ck[4]; vf[2] += ck[5];
      double vz_des = spline_vel_z(swing[k], p[L.p_swing_period], p[L.p_swing_height], p[L.p_swing_vel], p[L.p_swing_vel + 1]);
      ws.g[r0 + 5] = c * vf[0];
      ws.g[r0 + 6] = c * vf[1];
      ws.g[r0 + 7] = c * vf[2] + (1.0 - c) * (vf[2] - vz_des);
    }
    if (A.want_jac) {
      const int p0 = T.pos_foot[k];
      store_j(ws.J + p0, c);
      store_j(ws.J + p0 + 1, c * 2.0 * f[0]);
      store_j(ws.J + p0 + 2, c * 2.0 * f[1]);
      store_j(ws.J + p0 + 3, -c * mu2 * 2.0 * f[2]);
      store_j(ws.J + p0 + 4, 1.0 - c);
      store_j(ws.J + p0 + 5, 1.0 - c);
      store_j(ws.J + p0 + 6, 1.0 - c);
    }
  }
  if (M.has_ext && lane == M.nfeet) {
    for (int t = 0; t < 3; ++t) ws.g[T.row_ext + t] = ws.fx[3 * M.nfeet + t];
  }
  // ---- arm rows values (lane 5): base-relative linear velocity, world z
  if (T.row_arm >= 0 && lane == 5) {
    const double* ak = ws.arm;
    // relative spatial velocity wrt the base body (velocities are about the base origin)
    double Vb[6];
    matvec3(ws.Rb, ws.cv, Vb);
    matvec3(ws.Rb, ws.cv + 3, Vb + 3);
    double rel[6];
    for (int i = 0; i < 6; ++i) rel[i] = ak[3 + i] - Vb[i];
    double lin[3], linb[3], full[3];
    cross3(rel + 3, ak, lin);
    for (int i = 0; i < 3; ++i) lin[i] += rel[i];
    matTvec3(ws.Rb, lin, linb);
    cross3(ak + 6, ak, full);
    const double* des = p + L.p_arm_vel;
    ws.g[T.row_arm] = linb[0] - des[0];
    ws.g[T.row_arm + 1] = linb[1] - des[1];
    ws.g[T.row_arm + 2] = full[2] + ak[5] - des[2];
  }
  if (!A.want_jac || lane >= nv) return;
  // foot / arm velocity Jacobians (lane = column); centroidal_vel without base inputs emits them after phase E, when
  // the directions d v_b / d(column) are known, and only records the rows' coefficients on v_b here
  if (KIND == PLM_CENTROIDAL_VEL && NOBASE) node_phase_fjac<KIND, 1>(ws, A, st, lane);
  else node_phase_fjac<KIND, 0>(ws, A, st, lane);
}

// Constant Jacobian entries; lanes stride over the list.
PLM_HD void node_phase_consts(NodeWs& ws, const NodeArgs& A, int lane, int nlanes) {
  const PlmNodeType& T = *A.T;
  const double* contact = A.p + A.L->p_contact + 4 * A.node;
  for (int e = lane; e < T.nconst; e += nlanes) {
#if defined(__CUDA_ARCH__) && PLM_EMIT_HINTS
    PlmConstEntry ce;      // (the table is shared by every evaluation: kept in L1 like the lookup table)
    {
      int lo, hi;
      asm("ld.global.nc.L1::evict_last.v2.s32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(A.consts + e));
      ce.pos = lo; ce.code = (int16_t)(hi & 0xffff); ce.arg = (int16_t)(hi >> 16);
    }
#else
    const PlmConstEntry ce = A.consts[e];
#endif
    double v;
    switch (ce.code) {
      case 0: v = 1.0; break;
      case 1: v = -1.0; break;
      case 2: v = -A.dt; break;
      case 3: v = -A.M->total_mass; break;
      case 4: v = contact[ce.arg]; break;
      default: v = 1.0 - contact[ce.arg]; break;
    }
    store_j(ws.J + ce.pos, v);
  }
}

}  // namespace plm
