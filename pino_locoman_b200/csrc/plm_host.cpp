// Host table builder (see plm_host.h).  Row order and canonical forms follow optimization/ocp.py:103-198 and the
// setup_dynamics_constraints of each optimization/ocp_*.py; x / p layouts follow setup_variables / setup_parameters.
#include "plm_host.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>

namespace plm {

namespace {

struct Entry {
  int col;       // local column in [dx_i | u_i | dx_{i+1}]
  int kind;      // 0: lut source, 1: constant, 2: direct (written by a lane at a recorded position)
  int src, idx;  // kind 0: source block and index;  kind 1: code/arg
};

struct RowBuilder {
  std::vector<std::vector<Entry>> rows;
  int new_row() { rows.emplace_back(); return (int)rows.size() - 1; }
  void add(int row, int col, int kind, int src, int idx) { rows[row].push_back({col, kind, src, idx}); }
};

bool is_anc_or_self(const PlmModel& M, int a, int b) {   // body a is an ancestor-or-self of body b
  while (b >= 0) {
    if (a == b) return true;
    b = M.parent[b];
  }
  return false;
}

}  // namespace

static bool build_model(const plm_robot_desc& r, PlmModel& M, std::string& err) {
  memset(&M, 0, sizeof(M));
  if (r.nbody < 1 || r.nbody > PLM_MAXB) { err = "nbody out of range"; return false; }
  M.nbody = r.nbody;
  M.nv = 6 + r.nbody - 1;
  M.nq = 7 + r.nbody - 1;
  M.nj = r.nbody - 1;
  if (M.nv > PLM_MAXCOL) { err = "nv exceeds one warp"; return false; }
  M.nfeet = r.nfeet;
  M.has_ext = r.has_ext_force ? 1 : 0;
  M.ncontact = r.nfeet + M.has_ext;
  if (r.nfeet != 4 || M.ncontact > PLM_MAXC) { err = "expected 4 feet"; return false; }
  M.arm_body = r.arm_body;
  for (int i = 0; i < 3; ++i) M.arm_off[i] = r.arm_offset[i];
  M.gravity_z = 9.81;
  double total = 0;
  for (int b = 0; b < M.nbody; ++b) {
    M.parent[b] = r.parent[b];
    if (b == 0 ? r.parent[b] != -1 : (r.parent[b] < 0 || r.parent[b] >= b)) { err = "bodies must be in topological order"; return false; }
    const double* pl = r.placement + 12 * b;
    bool rot = false;
    for (int i = 0; i < 9; ++i) {
      M.place_R[b][i] = pl[i];
      double id = (i % 4 == 0) ? 1.0 : 0.0;
      if (fabs(pl[i] - id) > 0) rot = true;
    }
    M.has_rot[b] = rot;
    for (int i = 0; i < 3; ++i) M.place_p[b][i] = pl[9 + i];
    const double* ax = r.axis + 3 * b;
    double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    M.axtype[b] = 3;
    for (int i = 0; i < 3; ++i) M.axis[b][i] = (b > 0 && n > 0) ? ax[i] / n : 0.0;
    if (b > 0) {
      for (int i = 0; i < 3; ++i)
        if (M.axis[b][i] == 1.0 && M.axis[b][(i + 1) % 3] == 0.0 && M.axis[b][(i + 2) % 3] == 0.0) M.axtype[b] = i;
    }
    const double* in = r.inertia + 10 * b;
    M.mass[b] = in[0];
    total += in[0];
    for (int i = 0; i < 3; ++i) M.com[b][i] = in[1 + i];
    for (int i = 0; i < 6; ++i) M.Ic[b][i] = in[4 + i];
  }
  M.total_mass = total;
  for (int k = 0; k < M.ncontact; ++k) {
    M.contact_body[k] = r.contact_body[k];
    for (int i = 0; i < 3; ++i) M.contact_off[k][i] = r.contact_offset[3 * k + i];
    for (int k2 = 0; k2 < k; ++k2)
      if (M.contact_body[k2] == M.contact_body[k]) { err = "at most one contact frame per body"; return false; }
  }
  for (int d = 0; d < M.nv; ++d) {
    int body = d < 6 ? 0 : d - 5;
    M.col_body[d] = body;
    int chain[PLM_MAXB], len = 0;
    for (int b = body; b > 0; b = M.parent[b]) chain[len++] = b;
    if (len > PLM_MAXDEPTH) { err = "kinematic chain too deep"; return false; }
    M.chain_len[d] = len;
    for (int l = 0; l < len; ++l) M.chain[d][l] = chain[len - 1 - l];
    uint32_t mask = 0;
    for (int k = 0; k < M.ncontact; ++k)
      if (is_anc_or_self(M, body, M.contact_body[k])) mask |= 1u << k;
    M.col_contacts[d] = mask;
    if (M.arm_body >= 0 && is_anc_or_self(M, body, M.arm_body)) M.col_arm |= 1u << d;
  }
  int n = 0;
  for (int b = M.nbody - 1; b >= 1; --b) M.body_order[n++] = b;   // topological order => reverse is leaves first
  for (int j = 0; j < M.nj; ++j) {
    M.joint_pos_min[j] = r.joint_pos_min[j];
    M.joint_pos_max[j] = r.joint_pos_max[j];
    M.joint_vel_max[j] = r.joint_vel_max[j];
    M.joint_torque_max[j] = r.joint_torque_max[j];
  }
  for (int i = 0; i < M.nq; ++i) M.q0[i] = r.q0[i];
  return true;
}

bool build_tables(const plm_robot_desc& robot, const plm_ocp_desc& ocp, HostTables& out) {
  PlmModel& M = out.model;
  PlmLayout& L = out.layout;
  if (!build_model(robot, M, out.error)) return false;
  memset(&L, 0, sizeof(L));
  const int kind = ocp.dynamics;
  if (kind < 0 || kind > PLM_WHOLE_BODY_RNEA) { out.error = "Unknown dynamics type"; return false; }
  const int N = ocp.nodes;
  if (N < 2 || N > PLM_MAXNODES) { out.error = "nodes out of range"; return false; }
  const int nv = M.nv, nq = M.nq, nj = M.nj, nfeet = M.nfeet;
  const int nf = 3 * M.ncontact;
  L.dynamics = kind;
  L.nodes = N;
  L.tau_nodes = (kind == PLM_WHOLE_BODY_RNEA) ? std::min(ocp.tau_nodes, N) : 0;
  L.mu = ocp.mu;
  L.nf = nf;
  const bool cvel = kind == PLM_CENTROIDAL_VEL;
  L.nx = cvel ? 6 + nq : nq + nv;
  L.ndx = cvel ? 6 + nv : 2 * nv;
  const int ndx = L.ndx;
  const bool nobase = !ocp.include_base && (kind == PLM_CENTROIDAL_VEL || kind == PLM_CENTROIDAL_ACC || kind == PLM_WHOLE_BODY_ACC);
  L.nobase = nobase ? 1 : 0;
  const bool noacc = !ocp.include_acc && kind == PLM_WHOLE_BODY_RNEA;
  L.noacc = noacc ? 1 : 0;
  // leading input block: v / a (nv), tau_j (nj), the joint part only when the base part follows from the dynamics, or
  // nothing (whole_body_rnea with finite-difference accelerations)
  L.lead = noacc ? 0 : ((kind == PLM_WHOLE_BODY_ABA || nobase) ? nj : nv);
  L.f_idx = L.lead;
  L.tau_idx = L.lead + nf;
  // stage offsets
  std::vector<int> nu(N);
  for (int i = 0; i < N; ++i) nu[i] = L.lead + nf + ((kind == PLM_WHOLE_BODY_RNEA && i < L.tau_nodes) ? nj : 0);
  L.x_off[0] = 0;
  for (int i = 0; i < N; ++i) L.x_off[i + 1] = L.x_off[i] + ndx + nu[i];
  L.n = L.x_off[N] + ndx;
  // parameter offsets
  int off = 0;
  auto take = [&](int32_t& dst, int sz) { dst = off; off += sz; };
  take(L.p_x_init, L.nx); take(L.p_dt_min, 1); take(L.p_dt_max, 1); take(L.p_contact, 4 * N); take(L.p_swing, 4 * N);
  take(L.p_n_contacts, 1); take(L.p_swing_period, 1); take(L.p_swing_height, 1); take(L.p_swing_vel, 2);
  take(L.p_Q, ndx); take(L.p_R, nu[0]); take(L.p_base_vel, 6); take(L.p_ext_force, 3); take(L.p_arm_vel, 3);
  L.p_tau_prev = L.p_W = -1;
  if (kind == PLM_WHOLE_BODY_RNEA) { take(L.p_tau_prev, nj); take(L.p_W, nj); }
  L.np = off;

  // local column helpers
  const int dq0 = cvel ? 6 : 0;           // dq block inside dx
  auto col_dq = [&](int c) { return dq0 + c; };
  // dv (state) or U velocity; without base inputs the U block holds the joint part only (c >= 6)
  auto col_v = [&](int c, int nu_i) { (void)nu_i; return cvel ? ndx + c - (nobase ? 6 : 0) : nv + c; };
  auto col_lead = [&](int c) { return ndx + c; };
  auto col_leadj = [&](int c) { return ndx + c - (nobase ? 6 : 0); };   // acceleration column of velocity column c
  auto col_f = [&](int k, int t) { return ndx + L.f_idx + 3 * k + t; };
  auto col_tau = [&](int j) { return ndx + L.tau_idx + j; };
  auto col_next = [&](int c, int nu_i) { return ndx + nu_i + c; };

  // node types: (first-node skip?, torque rows?)
  const bool first_skip = !cvel;
  L.ntypes = 0;
  int type_of[2][2] = {{-1, -1}, {-1, -1}};
  std::vector<std::vector<std::vector<Entry>>> type_rows;
  for (int i = 0; i < N; ++i) {
    const int skip = (i == 0 && first_skip) ? 1 : 0;
    const int jr = (kind == PLM_WHOLE_BODY_RNEA && i < L.tau_nodes) ? 1 : 0;
    if (type_of[skip][jr] >= 0) { L.node_type[i] = type_of[skip][jr]; continue; }
    if (L.ntypes >= 4) { out.error = "too many node types"; return false; }
    const int t = L.ntypes++;
    type_of[skip][jr] = t;
    L.node_type[i] = t;
    PlmNodeType& T = L.types[t];
    memset(&T, 0, sizeof(T));
    T.nu = nu[i];
    T.joint_rows = jr;
    T.state_rows = !skip;
    T.row_int = T.row_dyn = T.row_tauj = T.row_taub = T.row_ext = T.row_arm = T.row_qj = T.row_vj = -1;
    const int nu_i = nu[i];
    // lut source offsets
    const int sizes[PLM_SRC_COUNT] = {nv * nv, nv * nv, nv * nv, nv * nf, nfeet * 3 * nv, nfeet * 3 * nv, 3 * nv, 3 * nv, 6 * nv, 6 * nf,
                                      36, nfeet * 3 * 6, 3 * 6, nv * nv};
    int lo = 0;
    for (int s = 0; s < PLM_SRC_COUNT; ++s) { T.src_off[s] = lo; lo += sizes[s]; }
    T.lut_size = lo;
    RowBuilder rb;
    auto q_cols_ok = [&](int c) { return c >= 3; };   // no row depends on the base translation increment
    // ---- 1. dynamics rows
    T.row_int = 0;
    if (cvel) {
      for (int r = 0; r < 6; ++r) {   // dh_next - (dh + h_dot dt)
        int row = rb.new_row();
        rb.add(row, r, 1, 1, 0);
        if (r >= 3) for (int c = 3; c < nv; ++c) rb.add(row, col_dq(c), 0, PLM_SRC_XQ, r * nv + c);
        for (int k = 0; k < M.ncontact; ++k)
          for (int tt = 0; tt < 3; ++tt)
            if (r >= 3 || tt == r) rb.add(row, col_f(k, tt), 0, PLM_SRC_XF, r * nf + 3 * k + tt);
        rb.add(row, col_next(r, nu_i), 1, 0, 0);
      }
      for (int c = 0; c < nv; ++c) {  // dq_next - (dq + v dt)
        int row = rb.new_row();
        if (nobase && c < 6) {
          // v_b = A_b^-1 (m h - A_j v_j) (dynamics_centroidal_vel.py:73-89): dense in dh, dq and v_j; the -1 of dq_c is
          // folded into the computed entry of column dq_c (c >= 3; the translation increments c < 3 have no other term)
          for (int k = 0; k < 6; ++k) rb.add(row, k, 0, PLM_SRC_IH, c * 6 + k);
          if (c < 3) rb.add(row, col_dq(c), 1, 1, 0);
          for (int d = 3; d < nv; ++d) rb.add(row, col_dq(d), 0, PLM_SRC_TQ, c * nv + d);
          for (int d = 6; d < nv; ++d) rb.add(row, col_v(d, nu_i), 0, PLM_SRC_TV, c * nv + d);
        } else {
          rb.add(row, col_dq(c), 1, 1, 0);
          rb.add(row, col_v(c, nu_i), 1, 2, 0);
        }
        rb.add(row, col_next(6 + c, nu_i), 1, 0, 0);
      }
      if (!nobase) {
        T.row_dyn = (int)rb.rows.size();
        for (int r = 0; r < 6; ++r) {   // A v - m h
          int row = rb.new_row();
          rb.add(row, r, 1, 3, 0);
          for (int c = 3; c < nv; ++c) rb.add(row, col_dq(c), 0, PLM_SRC_TQ, r * nv + c);
          for (int c = 0; c < nv; ++c) rb.add(row, col_v(c, nu_i), 0, PLM_SRC_TV, r * nv + c);
        }
      }
    } else {
      for (int c = 0; c < nv; ++c) {  // dq_next - (dq + v dt)
        int row = rb.new_row();
        rb.add(row, col_dq(c), 1, 1, 0);
        rb.add(row, col_v(c, nu_i), 1, 2, 0);
        rb.add(row, col_next(c, nu_i), 1, 0, 0);
      }
      for (int c = 0; c < nv && !noacc; ++c) {  // dv_next - (dv + a dt)
        int row = rb.new_row();
        if (kind == PLM_WHOLE_BODY_ABA) {
          for (int d = 3; d < nv; ++d) rb.add(row, col_dq(d), 0, PLM_SRC_TQ, c * nv + d);
          for (int d = 0; d < nv; ++d) rb.add(row, col_v(d, nu_i), 0, PLM_SRC_TV, c * nv + d);
          for (int j = 0; j < nj; ++j) rb.add(row, col_lead(j), 0, PLM_SRC_TA, c * nv + j);
          for (int k = 0; k < M.ncontact; ++k)
            for (int tt = 0; tt < 3; ++tt) rb.add(row, col_f(k, tt), 0, PLM_SRC_TF, c * nf + 3 * k + tt);
        } else if (nobase && c < 6) {
          // a_b = base_acc_dynamics(q, v, a_j, forces) (dynamics_centroidal_acc.py:43-82, dynamics_whole_body_acc.py:43-83):
          // dense in dq, dv, a_j and the forces; the -1 of dv_c is folded into the computed entry of column dv_c
          for (int d = 3; d < nv; ++d) rb.add(row, col_dq(d), 0, PLM_SRC_TQ, c * nv + d);
          for (int d = 0; d < nv; ++d) rb.add(row, col_v(d, nu_i), 0, PLM_SRC_TV, c * nv + d);
          for (int d = 6; d < nv; ++d) rb.add(row, col_leadj(d), 0, PLM_SRC_TA, c * nv + d);
          for (int k = 0; k < M.ncontact; ++k)
            for (int tt = 0; tt < 3; ++tt) rb.add(row, col_f(k, tt), 0, PLM_SRC_TF, c * nf + 3 * k + tt);
        } else {
          rb.add(row, col_v(c, nu_i), 1, 1, 0);
          rb.add(row, col_leadj(c), 1, 2, 0);
        }
        rb.add(row, col_next(nv + c, nu_i), 1, 0, 0);
      }
      if (!nobase && (kind == PLM_WHOLE_BODY_RNEA || kind == PLM_WHOLE_BODY_ACC || kind == PLM_CENTROIDAL_ACC)) {
        T.row_dyn = (int)rb.rows.size();
        const bool centroidal = kind == PLM_CENTROIDAL_ACC;
        const int nrow_t = 6 + (jr ? nj : 0);
        for (int r = 0; r < nrow_t; ++r) {
          if (r == 6) T.row_tauj = (int)rb.rows.size();
          int row = rb.new_row();
          const int rbody = M.col_body[r];
          for (int c = 0; c < nv; ++c) {
            const int cbody = M.col_body[c];
            const bool rel = is_anc_or_self(M, rbody, cbody) || is_anc_or_self(M, cbody, rbody);
            if (!rel) continue;
            if (q_cols_ok(c)) rb.add(row, col_dq(c), 0, PLM_SRC_TQ, r * nv + c);
            rb.add(row, col_v(c, nu_i), 0, PLM_SRC_TV, r * nv + c);
            if (noacc) rb.add(row, col_next(nv + c, nu_i), 0, PLM_SRC_TN, r * nv + c);      // a = (dv_next - dv) / dt
            else rb.add(row, col_lead(c), 0, PLM_SRC_TA, r * nv + c);
          }
          for (int k = 0; k < M.ncontact; ++k) {
            if (!is_anc_or_self(M, rbody, M.contact_body[k])) continue;
            for (int tt = 0; tt < 3; ++tt)
              if (!centroidal || r >= 3 || tt == r) rb.add(row, col_f(k, tt), 0, PLM_SRC_TF, r * nf + 3 * k + tt);
          }
          if (r >= 6) rb.add(row, col_tau(r - 6), 1, 1, 0);
        }
        if (jr) {
          T.row_taub = (int)rb.rows.size();
          for (int j = 0; j < nj; ++j) { int row = rb.new_row(); rb.add(row, col_tau(j), 1, 0, 0); }
        }
      }
    }
    // ---- 2. contact / swing rows per foot
    for (int k = 0; k < nfeet; ++k) {
      T.row_foot[k] = (int)rb.rows.size();
      int row = rb.new_row(); rb.add(row, col_f(k, 2), 2, k, 0);                       // c f_z >= 0
      row = rb.new_row(); for (int tt = 0; tt < 3; ++tt) rb.add(row, col_f(k, tt), 2, k, 1 + tt);   // cone
      for (int tt = 0; tt < 3; ++tt) { row = rb.new_row(); rb.add(row, col_f(k, tt), 2, k, 4 + tt); }  // (1-c) f = 0
      if (!skip) {
        for (int r = 0; r < 3; ++r) {   // c v_xy = 0 ; c v_z + (1-c)(v_z - v_des) = 0
          row = rb.new_row();
          if (cvel && nobase) {   // the foot velocity depends on v_b(dh, dq, v_j): every column
            for (int h6 = 0; h6 < 6; ++h6) rb.add(row, h6, 0, PLM_SRC_FH, (k * 3 + r) * 6 + h6);
            for (int c = 3; c < nv; ++c) rb.add(row, col_dq(c), 0, PLM_SRC_FQ, (k * 3 + r) * nv + c);
            for (int c = 6; c < nv; ++c) rb.add(row, col_v(c, nu_i), 0, PLM_SRC_FV, (k * 3 + r) * nv + c);
            continue;
          }
          for (int c = 0; c < nv; ++c) {
            if (!((M.col_contacts[c] >> k) & 1u)) continue;
            if (q_cols_ok(c)) rb.add(row, col_dq(c), 0, PLM_SRC_FQ, (k * 3 + r) * nv + c);
            rb.add(row, col_v(c, nu_i), 0, PLM_SRC_FV, (k * 3 + r) * nv + c);
          }
        }
      }
    }
    // ---- 3. external force rows
    if (M.has_ext) {
      T.row_ext = (int)rb.rows.size();
      for (int tt = 0; tt < 3; ++tt) { int row = rb.new_row(); rb.add(row, col_f(nfeet, tt), 1, 0, 0); }
    }
    // ---- 4. arm task and joint limits
    if (!skip) {
      if (M.arm_body >= 0) {
        T.row_arm = (int)rb.rows.size();
        for (int r = 0; r < 3; ++r) {
          int row = rb.new_row();
          if (cvel && nobase && r == 2) {   // world z of the frame velocity: depends on v_b(dh, dq, v_j)
            for (int h6 = 0; h6 < 6; ++h6) rb.add(row, h6, 0, PLM_SRC_AH, r * 6 + h6);
            for (int c = 3; c < nv; ++c) rb.add(row, col_dq(c), 0, PLM_SRC_AQ, r * nv + c);
            for (int c = 6; c < nv; ++c) rb.add(row, col_v(c, nu_i), 0, PLM_SRC_AV, r * nv + c);
            continue;
          }
          for (int c = 0; c < nv; ++c) {
            if (!((M.col_arm >> c) & 1u)) continue;
            if (r < 2 && M.col_body[c] == 0) continue;
            if (q_cols_ok(c)) rb.add(row, col_dq(c), 0, PLM_SRC_AQ, r * nv + c);
            rb.add(row, col_v(c, nu_i), 0, PLM_SRC_AV, r * nv + c);
          }
        }
      }
      T.row_qj = (int)rb.rows.size();
      for (int j = 0; j < nj; ++j) { int row = rb.new_row(); rb.add(row, col_dq(6 + j), 1, 0, 0); }
      T.row_vj = (int)rb.rows.size();
      for (int j = 0; j < nj; ++j) { int row = rb.new_row(); rb.add(row, col_v(6 + j, nu_i), 1, 0, 0); }
    }
    // ---- assign positions (row-major, columns ascending)
    T.nrows = (int)rb.rows.size();
    T.lut_off = (int)out.lut.size();
    out.lut.resize(out.lut.size() + T.lut_size, (int16_t)-1);
    T.const_off = (int)out.consts.size();
    int16_t* lut = out.lut.data() + T.lut_off;
    int pos = 0;
    for (int k = 0; k < 4; ++k) T.pos_foot[k] = -1;
    for (auto& row : rb.rows) {
      std::stable_sort(row.begin(), row.end(), [](const Entry& a, const Entry& b) { return a.col < b.col; });
      for (const Entry& e : row) {
        if (e.kind == 0) lut[T.src_off[e.src] + e.idx] = (int16_t)pos;
        else if (e.kind == 1) out.consts.push_back({pos, (int16_t)e.src, (int16_t)e.idx});
        else if (e.idx == 0) T.pos_foot[e.src] = pos;
        ++pos;
      }
    }
    if (pos > 32000) { out.error = "node block too large for int16 lut"; return false; }
    T.nnz = pos;
    T.nconst = (int)out.consts.size() - T.const_off;
    type_rows.push_back(rb.rows);
    {
      std::vector<std::vector<int>> rc;
      for (auto& row : rb.rows) { std::vector<int> cols; for (const Entry& e : row) cols.push_back(e.col); rc.push_back(cols); }
      out.type_rowcols.push_back(rc);
    }
  }
  // constant codes 4/5 (contact-scaled) are not used by row groups above: foot force rows are direct entries.

  // ---- global offsets and COO pattern
  L.row_off[0] = ndx;
  L.nnz_off[0] = ndx;
  L.max_rows = L.max_nnz = 0;
  for (int c = 0; c < ndx; ++c) { out.pat_rows.push_back(c); out.pat_cols.push_back(c); }   // DX_0 == 0
  for (int i = 0; i < N; ++i) {
    const PlmNodeType& T = L.types[L.node_type[i]];
    L.row_off[i + 1] = L.row_off[i] + T.nrows;
    L.nnz_off[i + 1] = L.nnz_off[i] + T.nnz;
    L.max_rows = std::max(L.max_rows, T.nrows);
    L.max_nnz = std::max(L.max_nnz, T.nnz);
    const auto& rows = type_rows[L.node_type[i]];
    for (size_t r = 0; r < rows.size(); ++r)
      for (const Entry& e : rows[r]) {
        out.pat_rows.push_back(L.row_off[i] + (int)r);
        out.pat_cols.push_back(L.x_off[i] + e.col);
      }
  }
  L.m = L.row_off[N];
  L.nnz = L.nnz_off[N];

  // ---- QP solver tables: per-type CSR/CSC of the node blocks, factor offsets, OSQP settings
  QpLayout& Q = out.qp;
  memset(&Q, 0, sizeof(Q));
  auto push = [&](const std::vector<int>& v) {
    int o = (int)out.qp_idx.size();
    for (int x : v) out.qp_idx.push_back((int16_t)x);
    return o;
  };
  for (int t = 0; t < L.ntypes; ++t) {
    const auto& rows = out.type_rowcols[t];
    const int s = ndx + L.types[t].nu, ncols = s + ndx;
    std::vector<int> rptr(1, 0), ccol;
    for (const auto& r : rows) { for (int c : r) ccol.push_back(c); rptr.push_back((int)ccol.size()); }
    std::vector<int> cptr(ncols + 1, 0), cpos(ccol.size()), crow(ccol.size());
    for (int c : ccol) cptr[c + 1]++;
    for (int c = 0; c < ncols; ++c) cptr[c + 1] += cptr[c];
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (size_t r = 0; r < rows.size(); ++r)
      for (int e = rptr[r]; e < rptr[r + 1]; ++e) { int c = ccol[e]; cpos[fill[c]] = e; crow[fill[c]] = (int)r; fill[c]++; }
    QpTypeIdx& I = Q.type[t];
    I.rptr = push(rptr); I.ccol = push(ccol); I.cptr = push(cptr); I.cpos = push(cpos); I.crow = push(crow);
    I.ncols = ncols; I.s = s;
    // structure the stage solver exploits: the first ndx rows are the integrator rows, each with exactly one entry in
    // DX_{i+1} (its last entry, at next-column r), and no other row touches DX_{i+1}; anything else takes the
    // general-coupling path (tables of the coupling rows below)
    std::vector<int> gc_rows, gc_rowq(rows.size(), -1);
    for (size_t r = 0; r < rows.size(); ++r) {
      int nnext = 0;
      for (int c : rows[r]) if (c >= s) nnext++;
      const bool integ = (int)r < ndx;
      if (integ ? (nnext != 1 || rows[r].back() != s + (int)r) : nnext != 0) Q.general_coupling = 1;
      if (nnext > 0) { gc_rowq[r] = (int)gc_rows.size(); gc_rows.push_back((int)r); }
    }
    I.ncoup = (int)gc_rows.size();
    I.gc_rows = push(gc_rows);
    I.gc_rowq = push(gc_rowq);
    Q.ncoup_max = std::max(Q.ncoup_max, I.ncoup);
  }
  {
    // flat CSR / CSC of the whole pattern (value order = CSR order)
    const int nnz = (int)out.pat_rows.size();
    std::vector<int> rptr(L.m + 1, 0), tptr(L.n + 1, 0), tsrc(nnz), rcol(nnz), trow(nnz);
    for (int e = 0; e < nnz; ++e) { rptr[out.pat_rows[e] + 1]++; tptr[out.pat_cols[e] + 1]++; rcol[e] = out.pat_cols[e]; }
    for (int r = 0; r < L.m; ++r) rptr[r + 1] += rptr[r];
    for (int j = 0; j < L.n; ++j) tptr[j + 1] += tptr[j];
    std::vector<int> fill(tptr.begin(), tptr.end() - 1);
    for (int e = 0; e < nnz; ++e) { const int j = out.pat_cols[e]; tsrc[fill[j]] = e; trow[fill[j]] = out.pat_rows[e]; fill[j]++; }
    auto push32 = [&](const std::vector<int>& v) { int o = (int)out.qp_idx32.size(); for (int x : v) out.qp_idx32.push_back(x); return o; };
    Q.f_rptr = push32(rptr); Q.f_tptr = push32(tptr); Q.f_tsrc = push32(tsrc);
    Q.f_rcol = push(rcol); Q.f_trow = push(trow);
    std::vector<int> rperm(L.m), cperm(L.n);
    for (int r = 0; r < L.m; ++r) rperm[r] = r;
    for (int j = 0; j < L.n; ++j) cperm[j] = j;
    std::stable_sort(rperm.begin(), rperm.end(), [&](int a, int b) { return rptr[a + 1] - rptr[a] > rptr[b + 1] - rptr[b]; });
    std::stable_sort(cperm.begin(), cperm.end(), [&](int a, int b) { return tptr[a + 1] - tptr[a] > tptr[b + 1] - tptr[b]; });
    Q.f_rperm = push(rperm); Q.f_cperm = push(cperm);
    // sliced-ELL copies for the ADMM products: items (rows or columns) in order of decreasing length, 32 per slice, slot
    // (j, lane) of a slice at base + 32 j + lane; src = CSR value position (-1 padding), ind = gathered index
    auto build_ell = [&](const std::vector<int>& perm, const std::vector<int>& ptr, const std::vector<int>* srcmap, const std::vector<int>& indmap,
                         int32_t& f_base, int32_t& f_src, int32_t& f_ind, int32_t& nsl, int32_t& total) {
      const int nitems = (int)perm.size();
      nsl = (nitems + 31) / 32;
      std::vector<int> base(1, 0), src, ind;
      for (int sl = 0; sl < nsl; ++sl) {
        const int width = ptr[perm[32 * sl] + 1] - ptr[perm[32 * sl]];
        for (int j = 0; j < width; ++j)
          for (int lane = 0; lane < 32; ++lane) {
            const int it = 32 * sl + lane;
            int e = -1;
            if (it < nitems && j < ptr[perm[it] + 1] - ptr[perm[it]]) e = ptr[perm[it]] + j;
            src.push_back(e < 0 ? -1 : (srcmap ? (*srcmap)[e] : e));
            ind.push_back(e < 0 ? 0 : indmap[e]);
          }
        base.push_back((int)src.size());
      }
      total = (int)src.size();
      f_base = push32(base); f_src = push32(src); f_ind = push(ind);
    };
    build_ell(rperm, rptr, nullptr, rcol, Q.f_rell_base, Q.f_rell_src, Q.f_rell_ind, Q.n_rslices, Q.rell_total);
    build_ell(cperm, tptr, &tsrc, trow, Q.f_cell_base, Q.f_cell_src, Q.f_cell_ind, Q.n_cslices, Q.cell_total);
  }
  Q.sparse_coupling = Q.general_coupling ? 0 : 1;
  for (int t = 0; t < L.ntypes && !Q.general_coupling; ++t)
    for (int r = 0; r < ndx; ++r)
      if ((int)out.type_rowcols[t][r].size() - 1 > 4) Q.sparse_coupling = 0;
  Q.smax = 0;
  int fo = 0;
  for (int i = 0; i <= N; ++i) {
    const int s = (i < N) ? ndx + nu[i] : ndx;
    Q.fac_off[i] = fo;
    fo += (plm_sinv_rows(s) * s + 1) & ~1;   // cyclic-diagonal array of S_i^-1 (plm_qp_types.h); blocks start 16-byte aligned (bulk copies)
    Q.smax = std::max(Q.smax, s);
  }
  Q.fac_off[N + 1] = fo;
  for (int i = 0; i < N; ++i) {         // B_i = S_i^-1 G_i^T: ndx columns of sp doubles
    Q.bk_off[i] = fo;
    fo += ndx * ((ndx + nu[i] + 1) & ~1);
  }
  Q.bk_off[N] = fo;
  Q.fac_total = fo;
  {
    // panel schedules: forward sweep stages 0..N (row panels of the cyclic-diagonal array of S_i^-1), backward sweep
    // stages N-1..0 (column panels of B_i), panels of at most `capacity` doubles.  Two schedules: PLM_PANEL_DOUBLES panels
    // for the throughput kernel (four CTAs per SM), whole stages for the latency kernel (one CTA per SM, shared memory to spare).
    bool sched_ok = true;
    auto build = [&](int capacity, int32_t& f_sched, int32_t& n_sched, int32_t& panel_doubles) {
      std::vector<int> sched;
      int maxlen = 0;
      auto add_stage = [&](int i, int dir) {
        const int s = (i < N) ? ndx + nu[i] : ndx;
        const int rows = plm_sinv_rows(s);
        const int maxrows = std::max(1, (capacity - 2) / s);      // (- 2: the copy starts / ends on 16-byte boundaries)
        const int npan = (rows + maxrows - 1) / maxrows;
        for (int k = 0; k < npan; ++k) {
          const int r0 = (int)((long long)rows * k / npan), r1 = (int)((long long)rows * (k + 1) / npan);
          const int o0 = r0 * s, o1 = r1 * s;
          const int start = o0 & ~1;                       // 16-byte aligned start (stage blocks start even)
          const int len = ((o1 - start) + 1) & ~1;
          maxlen = std::max(maxlen, len);
          sched.push_back(Q.fac_off[i] + start); sched.push_back(len); sched.push_back(r0); sched.push_back(r1);
          sched.push_back(i);
          // (bits 8.. of the flags: size of the previous stage, for the coupling step of the first panel)
          sched.push_back(dir | ((k == 0) << 1) | ((k + 1 == npan) << 2) | ((i > 0 ? ndx + nu[i - 1] : 0) << 8));
          sched.push_back(start);
          sched.push_back(s | (L.x_off[i] << 8));
        }
      };
      // backward stage i: x_i = tv_i - B_i x_{i+1}[0:ndx]; equal column panels
      auto add_back = [&](int i) {
        const int s = ndx + nu[i], sp = (s + 1) & ~1;
        const int maxcols = std::max(1, capacity / sp);
        const int npan = (ndx + maxcols - 1) / maxcols;
        for (int k = 0; k < npan; ++k) {
          const int j0 = (int)((long long)ndx * k / npan), j1 = (int)((long long)ndx * (k + 1) / npan);
          const int len = (j1 - j0) * sp;
          maxlen = std::max(maxlen, len);
          sched.push_back(Q.bk_off[i] + j0 * sp); sched.push_back(len); sched.push_back(j0); sched.push_back(j1);
          sched.push_back(i);
          sched.push_back(1 | ((k == 0) << 1) | ((k + 1 == npan) << 2));
          sched.push_back(sp);
          sched.push_back(s | (L.x_off[i] << 8));
        }
      };
      // dense integrator rows: the coupling block of node i - 1 travels through the ring ahead of stage i's panels
      auto add_coupling = [&](int i) {
        const int gld = Q.gdense_ld;
        const int maxrows = std::max(1, capacity / gld);
        const int npan = (ndx + maxrows - 1) / maxrows;
        for (int k = 0; k < npan; ++k) {
          const int c0 = (int)((long long)ndx * k / npan), c1 = (int)((long long)ndx * (k + 1) / npan);
          const int len = (c1 - c0) * gld;
          maxlen = std::max(maxlen, len);
          sched.push_back(((i - 1) * ndx + c0) * gld); sched.push_back(len); sched.push_back(c0); sched.push_back(c1);
          sched.push_back(i);
          sched.push_back(8 | ((k == 0) << 4) | ((ndx + nu[i - 1]) << 8));
          sched.push_back(gld);
          sched.push_back((i < N ? ndx + nu[i] : ndx) | (L.x_off[i] << 8));
        }
      };
      for (int i = 0; i <= N; ++i) {
        if (Q.gdense_ld > 0 && i > 0) add_coupling(i);
        add_stage(i, 0);
      }
      for (int i = N - 1; i >= 0; --i) add_back(i);         // x_N = S_N^-1 r_N needs no backward work
      n_sched = (int)sched.size() / PLM_SCHED_INTS;
      // the ADMM kernel packs an entry into 16 bytes (plm_qp.cu): field widths
      for (int k = 0; k < n_sched; ++k) {
        const int* e = sched.data() + k * PLM_SCHED_INTS;
        if (e[1] > 0xffff || e[2] > 255 || e[3] > 255 || e[4] > 255 || (e[5] >> 8) > 255 || e[6] > 0xffff || (e[7] >> 8) > 0xffff)
          sched_ok = false;
      }
      while (out.qp_idx32.size() % 4) out.qp_idx32.push_back(0);     // schedule entries are read as two 16-byte words
      f_sched = (int)out.qp_idx32.size();
      for (int v : sched) out.qp_idx32.push_back(v);
      panel_doubles = (maxlen + 1) & ~1;
    };
    int cap = PLM_PANEL_DOUBLES;
    if (const char* e = getenv("PLM_PANEL_DOUBLES")) cap = std::max(256, atoi(e));      // tuning hook (tools/ab_libs.sh)
    Q.gdense_ld = (!Q.sparse_coupling && !Q.general_coupling) ? ((Q.smax + 1) & ~1) : 0;
    build(cap, Q.f_sched, Q.n_sched, Q.panel_doubles);
    build(plm_sinv_rows(Q.smax) * Q.smax + 4, Q.f_sched_lat, Q.n_sched_lat, Q.panel_doubles_lat);
    if (!sched_ok || Q.smax > 255 || N > 254) { out.error = "problem too large for the packed ADMM schedule (stage size / node count / offsets)"; return false; }
    Q.g_doubles = (5 * ndx + 1) & ~1;                              // compact coupling block of one stage: <= 4 entries per integrator row + their columns (packed)
  }
  Q.max_iter = ocp.osqp_max_iter; Q.check_termination = ocp.osqp_check_termination; Q.scaling = ocp.osqp_scaling;
  Q.rho = ocp.osqp_rho; Q.sigma = ocp.osqp_sigma; Q.alpha = ocp.osqp_alpha;
  Q.eps_abs = ocp.osqp_eps_abs; Q.eps_rel = ocp.osqp_eps_rel;
  Q.eps_prim_inf = ocp.osqp_eps_prim_inf; Q.eps_dual_inf = ocp.osqp_eps_dual_inf;
  return true;
}

}  // namespace plm
