// Plain-old-data tables shared by the host table builder, the CUDA kernels and the host emulation
// used by the CPU tests.  Naming follows the reference's domain: bodies/joints of the kinematic tree
// (utils/robot.py), shooting nodes and stages of the OCP (optimization/ocp.py), contacts = foot frames
// plus the optional external-force frame.
#pragma once
#include <stdint.h>

#define PLM_MAXB 24      // movable joints (bodies) incl. the free-flyer root
#define PLM_MAXCOL 32    // nv (one warp lane per velocity column)
#define PLM_MAXC 6       // contact frames: 4 feet + external-force frame (+ spare)
#define PLM_MAXDEPTH 8   // revolute joints between the root and a leaf
#define PLM_MAXNODES 64
#define PLM_NODE_WARPS 8   // upper bound of warps (= node evaluations) per CTA of the node kernel

enum PlmDynamics {
  PLM_CENTROIDAL_VEL = 0,
  PLM_CENTROIDAL_ACC = 1,
  PLM_WHOLE_BODY_ACC = 2,
  PLM_WHOLE_BODY_ABA = 3,
  PLM_WHOLE_BODY_RNEA = 4,
};

// Kinematic-tree tables (one per robot), resident in HBM and staged to shared memory by the kernels.
struct PlmModel {
  int32_t nbody, nq, nv, nj, ncontact, nfeet;
  int32_t has_ext;            // contact[nfeet] is the external-force frame
  int32_t arm_body;           // parent body of the arm end-effector frame, -1 if none
  double arm_off[3];
  double total_mass;
  double gravity_z;           // 9.81 (model gravity is (0,0,-9.81))
  int32_t parent[PLM_MAXB];   // parent body, -1 for the root
  int32_t axtype[PLM_MAXB];   // 0/1/2: revolute about x/y/z of the joint frame, 3: general axis
  int32_t has_rot[PLM_MAXB];  // joint placement has a non-identity rotation
  double axis[PLM_MAXB][3];
  double place_p[PLM_MAXB][3];
  double place_R[PLM_MAXB][9];   // row-major
  double mass[PLM_MAXB];
  double com[PLM_MAXB][3];       // in the joint frame
  double Ic[PLM_MAXB][6];        // xx xy xz yy yz zz about the com, joint-frame axes
  int32_t contact_body[PLM_MAXC];
  double contact_off[PLM_MAXC][3];
  // per velocity column (lane)
  int32_t col_body[PLM_MAXCOL];
  int32_t chain_len[PLM_MAXCOL];                  // revolute bodies from the root's child down to col_body
  int32_t chain[PLM_MAXCOL][PLM_MAXDEPTH];
  uint32_t col_contacts[PLM_MAXCOL];              // bit k: contact k's body is in the subtree of the column's joint
  uint32_t col_arm;                               // bit d: arm frame is in the subtree of column d
  int32_t body_order[PLM_MAXB];                   // non-root bodies, leaves first (composite accumulation order)
  // limits (utils/robot.py:52-55,65-68,91-118) and nominal configuration
  double joint_pos_min[PLM_MAXCOL], joint_pos_max[PLM_MAXCOL], joint_vel_max[PLM_MAXCOL], joint_torque_max[PLM_MAXCOL];
  double q0[PLM_MAXCOL + 1];
};

// Source blocks of per-node derivative entries (lanes write value -> lut[src] position in the node's J block).
enum PlmSrc {
  PLM_SRC_TQ = 0,   // [nv][nv]  d(dynamics rows)/d dq
  PLM_SRC_TV,       // [nv][nv]  d/d v-like variable (dv, or U velocity for centroidal_vel)
  PLM_SRC_TA,       // [nv][nv]  d/d a (or tau_j for ABA, column index = joint)
  PLM_SRC_TF,       // [nv][nf]  d/d forces
  PLM_SRC_FQ,       // [nfeet][3][nv] foot velocity rows d/d dq
  PLM_SRC_FV,       // [nfeet][3][nv]
  PLM_SRC_AQ,       // [3][nv] arm rows
  PLM_SRC_AV,       // [3][nv]
  PLM_SRC_XQ,       // [6][nv] centroidal_vel: d(h_dot)/d dq
  PLM_SRC_XF,       // [6][nf] centroidal_vel: d(h_dot)/d forces
  PLM_SRC_IH,       // [6][6]  centroidal_vel without base inputs: d(base dq-integrator rows)/d dh
  PLM_SRC_FH,       // [nfeet][3][6]  "  : d(foot velocity rows)/d dh
  PLM_SRC_AH,       // [3][6]         "  : d(arm rows)/d dh
  PLM_SRC_TN,       // [nv][nv] whole_body_rnea without acceleration inputs: d(dynamics rows)/d dv_{i+1}
  PLM_SRC_COUNT
};

// A constant (x-independent) Jacobian entry of a node block.
struct alignas(8) PlmConstEntry {      // one 8-byte load per entry
  int32_t pos;     // position in the node's J block
  int16_t code;    // 0: +1, 1: -1, 2: -dt, 3: -mass, 4: contact k (c_k), 5: 1-c_k   (k in arg)
  int16_t arg;
};

// Per node-type tables (a node type fixes the row list: first-node skip, torque rows).
struct PlmNodeType {
  int32_t nrows, nnz;
  int32_t nu;                     // input size at this node
  int32_t joint_rows;             // RNEA: torque rows present
  int32_t state_rows;             // foot-velocity / arm / joint-limit rows present (not the first-node skip)
  int32_t src_off[PLM_SRC_COUNT]; // offsets into lut
  int32_t lut_size;
  int32_t lut_off;                // offset of this type's lut in the global int16 pool
  int32_t nconst, const_off;      // constant entries in the global PlmConstEntry pool
  // row offsets of row groups inside the node block (-1 if absent)
  int32_t row_int;      // integrator rows
  int32_t row_dyn;      // dynamics rows (rnea base / gaps)
  int32_t row_tauj;     // rnea joint rows
  int32_t row_taub;     // torque bound rows
  int32_t row_foot[4];  // first row of each foot's group (fz, cone, 3x zero-force, [vxy(2), vz])
  int32_t row_ext;
  int32_t row_arm;
  int32_t row_qj;
  int32_t row_vj;
  // positions (in the node's J block) of the direct entries of foot k: fz(1), cone(3), zero-force(3)
  int32_t pos_foot[4];
  int32_t pos_ext, pos_qj, pos_vj, pos_taub;
};

// Whole-problem layout (per OCP formulation), shared by every instance of the batch.
struct PlmLayout {
  int32_t dynamics, nodes, tau_nodes;
  int32_t nobase;                      // include_base = False: inputs without the base part, no dynamics-gap rows
  int32_t noacc;                       // include_acc = False (whole_body_rnea): a = (v_{i+1} - v_i) / dt, no dv integrator rows
  int32_t nx, ndx, n, m, np, nnz;
  int32_t nf;
  int32_t f_idx, tau_idx, lead;    // offsets inside U_i: forces, torques; size of the leading block (a / v / tau_j)
  double mu;
  int32_t x_off[PLM_MAXNODES + 1];     // stage offsets in x
  int32_t row_off[PLM_MAXNODES + 1];   // first g row of node i (row_off[0] = ndx)
  int32_t nnz_off[PLM_MAXNODES + 1];   // first J value of node i (nnz_off[0] = ndx)
  int32_t node_type[PLM_MAXNODES];
  int32_t ntypes;
  PlmNodeType types[4];
  int32_t max_rows, max_nnz;
  // parameter vector offsets (creation order of optimization/ocp.py:54-69, ocp_whole_body_rnea.py:88-89)
  int32_t p_x_init, p_dt_min, p_dt_max, p_contact, p_swing, p_n_contacts, p_swing_period, p_swing_height,
      p_swing_vel, p_Q, p_R, p_base_vel, p_ext_force, p_arm_vel, p_tau_prev, p_W;
};

// Device pointers to the tables of one handle (passed by value to the kernels).
#ifdef __cplusplus
namespace plm {
struct DeviceTables {
  const PlmModel* model;
  const PlmLayout* layout;
  const int16_t* lut;
  const PlmConstEntry* consts;
};
}  // namespace plm
#endif
