/* pino_locoman_b200 -- C ABI of the B200-native SQP inner loop (libpinolocoman_b200.so).
 *
 * Drop-in boundary for the hot path of lukasmolnar/pino-locoman: the casadi Functions built in
 * optimization/ocp.py:283-290 (sqp_data / f_data / g_data / hess_data), the Dynamics* Functions of
 * dynamics/*.py, the OSQP calls at optimization/ocp.py:312-313,395,401 and the Armijo line search at
 * optimization/ocp.py:430-480 -- batched over independent MPC instances.  The reference reaches
 * compiled code through casadi.external("sqp_data", lib) (optimization/ocp.py:299-301); this header
 * is the batched, device-pointer equivalent of that mechanism.
 *
 * Conventions
 *  - every `double*` / `int32_t*` named d_* is a DEVICE pointer owned by the caller (e.g. a torch CUDA
 *    tensor's data_ptr); everything else is host memory.  All arithmetic is FP64.
 *  - batched arrays are instance-major and contiguous: x is [batch][n], p is [batch][np], g is [batch][m],
 *    J values are [batch][nnz] in the fixed pattern returned by plm_jac_pattern().
 *  - x = [DX_0, U_0, DX_1, U_1, ..., DX_N] (e.g. optimization/ocp_whole_body_rnea.py:65-84, :293-324);
 *    p = the Opti parameters in creation order (optimization/ocp.py:54-69, ocp_whole_body_rnea.py:88-89),
 *    matrices column-major.
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it unless stated otherwise.
 *  - every function returns 0 on success, non-zero on error; plm_last_error() gives the message.
 *  - one handle per (GPU, stream); handles share no mutable state.
 */
#ifndef PINO_LOCOMAN_B200_H
#define PINO_LOCOMAN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plm_handle plm_handle;

/* Dynamics formulation: keys of ocp_factory.py:9-15 / ocp_args.py:2-19. */
enum { PLM_DYN_CENTROIDAL_VEL = 0, PLM_DYN_CENTROIDAL_ACC = 1, PLM_DYN_WHOLE_BODY_ACC = 2,
       PLM_DYN_WHOLE_BODY_ABA = 3, PLM_DYN_WHOLE_BODY_RNEA = 4 };

/* Kinematic-tree tables produced by the host loader (replaces pin.Model built in utils/robot.py:14-35).
 * Body 0 is the free-flyer root; bodies 1.. are revolute joints in pinocchio joint order, velocity
 * column of body b>0 is b+5.  Inertias are those of pinocchio's model.inertias (fixed links merged). */
typedef struct {
  int32_t nbody;                 /* movable joints incl. root */
  const int32_t* parent;         /* [nbody] parent body, -1 for the root */
  const double* placement;       /* [nbody][12] joint placement in the parent joint frame: R row-major (9), p (3) */
  const double* axis;            /* [nbody][3] revolute axis in the joint frame (ignored for the root) */
  const double* inertia;         /* [nbody][10] mass, com (3), I_c xx xy xz yy yz zz (joint frame, about the com) */
  int32_t nfeet;                 /* 4, order FR FL RR RL (utils/gait_sequence.py:7) */
  int32_t has_ext_force;         /* robot.ext_force_frame is set (utils/robot.py:70-76,98) */
  const int32_t* contact_body;   /* [nfeet + has_ext_force] parent body of each contact frame */
  const double* contact_offset;  /* [nfeet + has_ext_force][3] frame translation in the parent joint frame */
  int32_t arm_body;              /* parent body of robot.arm_ee_frame, -1 if none (utils/robot.py:99) */
  double arm_offset[3];
  const double* joint_pos_min;   /* [nj] utils/robot.py:52-55,65-68,91-118 */
  const double* joint_pos_max;
  const double* joint_vel_max;
  const double* joint_torque_max;
  const double* q0;              /* [nq] reference configuration */
} plm_robot_desc;

typedef struct {
  int32_t dynamics;              /* PLM_DYN_* */
  int32_t nodes;                 /* N */
  int32_t tau_nodes;             /* whole_body_rnea only (ocp_args.py:16) */
  double mu;                     /* friction coefficient, 0.7 (optimization/ocp.py:103) */
  /* OSQP settings, optimization/ocp.py:267-273 + osqp defaults */
  int32_t osqp_max_iter;         /* 100 */
  int32_t osqp_check_termination;/* 25 */
  int32_t osqp_scaling;          /* 10 */
  double osqp_rho, osqp_sigma, osqp_alpha;            /* 2e-2, 1e-6, 1.4 */
  double osqp_eps_abs, osqp_eps_rel;                  /* 1e-3, 1e-3 */
  double osqp_eps_prim_inf, osqp_eps_dual_inf;        /* 1e-4, 1e-4 */
  /* include_base (ocp_args.py:3-11; centroidal_vel / centroidal_acc / whole_body_acc): 1 = base velocity / acceleration
   * among the inputs + the 6 dynamics-gap rows (the reference's OCP_ARGS default); 0 = inputs without the base part,
   * v_b = base_vel_dynamics(h, q, v_j) / a_b = base_acc_dynamics(q, v, a_j, forces) substituted and no gap rows
   * (ocp_centroidal_vel.py:19-23,104-120; ocp_centroidal_acc.py:19-23,108-140; ocp_whole_body_acc.py:20-24,109-141). */
  int32_t include_base;
  /* include_acc (ocp_args.py:17; whole_body_rnea): 1 = accelerations among the inputs + the dv integrator rows (the
   * reference's OCP_ARGS default); 0 = no acceleration inputs, a_i = (v_{i+1} - v_i) / dt_i substituted and the dv
   * integrator rows dropped (ocp_whole_body_rnea.py:21-25,156,183-191): the RNEA rows then touch dv_{i+1}. */
  int32_t include_acc;
} plm_ocp_desc;

typedef struct {
  int32_t nq, nv, nj, nf;
  int32_t nx, ndx;               /* state / state-increment sizes */
  int32_t n, m, np, nnz;         /* decision variables, constraint rows, parameters, nnz(J_g) */
  int32_t nodes;
  int32_t kkt_factor_doubles;    /* stored size of the per-instance QP factor (doubles) */
} plm_dims;

void plm_fill_default_ocp_desc(plm_ocp_desc* desc, int32_t dynamics, int32_t nodes);
/* sizeof(plm_robot_desc), sizeof(plm_ocp_desc), sizeof(plm_dims): lets a binding verify its struct images. */
void plm_abi_struct_sizes(int32_t out[3]);

/* Create / destroy.  max_batch sizes the workspaces owned by the handle (QP factor, ADMM iterates).
 * max_batch == 0 gives a layout-only handle (dims / offsets / pattern queries work without a GPU; every
 * compute entry point fails).  With max_batch > 0 a CUDA device is required: there is no CPU path. */
int plm_create(const plm_robot_desc* robot, const plm_ocp_desc* ocp, int32_t max_batch, plm_handle** out);
void plm_destroy(plm_handle* h);
const char* plm_last_error(const plm_handle* h);
int plm_get_dims(const plm_handle* h, plm_dims* dims);

/* Layout queries (host arrays). */
int plm_stage_offsets(const plm_handle* h, int32_t* x_off /*[N+2]*/, int32_t* nu /*[N]*/, int32_t* row_off /*[N+2]*/);
int plm_param_offsets(const plm_handle* h, int32_t* off /*[16]*/);
/* COO pattern of J_g in the value order used by every J array: rows[nnz], cols[nnz]. */
int plm_jac_pattern(const plm_handle* h, int32_t* rows, int32_t* cols);

/* ---- casadi-Function equivalents (optimization/ocp.py:287-290), batched ----------------------- */
/* sqp_data(x,p) -> [grad_f, J_g, g, lbg, ubg] */
int plm_sqp_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch,
                 double* d_grad_f, double* d_J, double* d_g, double* d_lbg, double* d_ubg, void* stream);
/* g_data(x,p) -> [g, lbg, ubg]   (d_lbg / d_ubg may be NULL) */
int plm_g_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch,
               double* d_g, double* d_lbg, double* d_ubg, void* stream);
/* f_data(x,p) -> [f, grad_f]     (d_grad_f may be NULL) */
int plm_f_data(plm_handle* h, const double* d_x, const double* d_p, int32_t batch,
               double* d_f, double* d_grad_f, void* stream);
/* diag(hess_data(x,p)) -> [batch][n]  (constant diagonal, optimization/ocp.py:293-296) */
int plm_hess_diag(plm_handle* h, const double* d_p, int32_t batch, double* d_hess, void* stream);

/* ---- Dynamics* Functions (dynamics/*.py), batched; Jacobian outputs optional (NULL) ------------- */
/* integrate(x, dx) -> x_next ; difference(x0, x1) -> dx   (state layout of the handle's dynamics) */
int plm_state_integrate(plm_handle* h, const double* d_x, const double* d_dx, int32_t batch, double* d_x_next, void* stream);
int plm_state_difference(plm_handle* h, const double* d_x0, const double* d_x1, int32_t batch, double* d_dx, void* stream);
/* rnea_dyn(q, v, a, forces) -> tau_rnea [nv];  d_jac: [batch][nv][3*nv + nf] = d tau / d(dq_tangent, v, a, forces) */
int plm_rnea_dyn(plm_handle* h, const double* d_q, const double* d_v, const double* d_a, const double* d_forces,
                 int32_t batch, double* d_tau, double* d_jac, void* stream);
/* aba_dyn(q, v, tau_j, forces) -> a [nv];  d_jac: [batch][nv][2*nv + nj + nf] */
int plm_aba_dyn(plm_handle* h, const double* d_q, const double* d_v, const double* d_tau_j, const double* d_forces,
                int32_t batch, double* d_a, double* d_jac, void* stream);
/* centroidal_acc / whole_body_acc dyn_gaps(q, v, a, forces) -> gaps [6]; d_jac [batch][6][3*nv + nf] */
int plm_dyn_gaps(plm_handle* h, int32_t dynamics, const double* d_q, const double* d_v, const double* d_a,
                 const double* d_forces, int32_t batch, double* d_gaps, double* d_jac, void* stream);
/* centroidal_vel: dyn_gaps(h, q, v) -> gaps [6] ; com_dyn(q, forces) -> dh [6] */
int plm_centroidal_vel_gaps(plm_handle* h, const double* d_h, const double* d_q, const double* d_v, int32_t batch,
                            double* d_gaps, void* stream);
int plm_com_dyn(plm_handle* h, const double* d_q, const double* d_forces, int32_t batch, double* d_dh, void* stream);
/* base_acc(q, v, a_j, forces) -> a_b [6] for dynamics = centroidal_acc (dynamics_centroidal_acc.py:43-82, also
 * dynamics_centroidal_vel.py:91-134) or whole_body_acc (dynamics_whole_body_acc.py:43-83) with d_h = NULL;
 * base_vel(h, q, v_j) -> v_b [6] for dynamics = centroidal_vel (dynamics_centroidal_vel.py:73-89) with d_v = d_forces = NULL. */
int plm_base_solve(plm_handle* h, int32_t dynamics, const double* d_h, const double* d_q, const double* d_v,
                   const double* d_lead_j, const double* d_forces, int32_t batch, double* d_base, void* stream);
/* frame_vel(q, v) -> linear part vel[:3] (the components the OCP rows use, optimization/ocp.py:143-180) of
 * dynamics/dynamics.py:77-118: foot frame `contact` in 0..nfeet-1 with relative_to_base = 0 (LOCAL_WORLD_ALIGNED),
 * or the arm end-effector frame (contact = -1) with relative_to_base = 1 (x, y in base axes, z in the world).
 * d_vel is [batch][3]. */
int plm_frame_vel(plm_handle* h, int32_t contact, int32_t relative_to_base, const double* d_q, const double* d_v,
                  int32_t batch, double* d_vel, void* stream);
/* frame_pos(q) -> pos [batch][3] and frame_vel(q, v) -> vel [batch][6] of ANY frame (dynamics/dynamics.py:67-118): the
 * frame is given by its parent body (0 = free-flyer root, b > 0 = revolute joint b) and its placement in that body's
 * joint frame (placement12 = R row-major (9) | p (3), host memory).  vel = getFrameVelocity(LOCAL_WORLD_ALIGNED) =
 * [linear; angular]; with relative_to_base != 0 the reference's base-relative variant (dynamics.py:86-113, z components
 * kept in the world frame), the base frame given the same way (base_placement12 NULL = identity).  d_pos or d_vel may
 * be NULL; d_v may be NULL when d_vel is. */
int plm_frame_kinematics(plm_handle* h, int32_t body, const double* placement12, int32_t base_body, const double* base_placement12,
                         int32_t relative_to_base, const double* d_q, const double* d_v, int32_t batch, double* d_pos, double* d_vel,
                         void* stream);

/* ---- OSQP equivalents (optimization/ocp.py:312-313,395,401) -------------------------------------- */
/* setup(P, q=1, A=ones(pattern), l=-1, u=1) of optimization/ocp.py:305-313: zero the persistent ADMM iterates
 * (x, z, y) of the first `batch` instances and record the setup-time row scaling that osqp still has in force
 * when the first update() classifies the constraint rows.  d_hess is the diagonal P [batch][n]. */
int plm_qp_setup(plm_handle* h, int32_t batch, const double* d_hess, void* stream);
/* update(q=, Ax=, l=, u=): scale (Ruiz), classify rows, build and factor the per-instance stage-structured system.
 * d_hess is the diagonal P. */
int plm_qp_update(plm_handle* h, int32_t batch, const double* d_hess, const double* d_q, const double* d_J,
                  const double* d_l, const double* d_u, void* stream);
/* solve().x : ADMM from the persistent iterates; d_iters / d_status are int32 [batch] (may be NULL).
 * status: osqp codes (1 solved, 2 solved inaccurate, -2 max iterations, -3/3 primal infeasible (in)accurate,
 * -4/4 dual infeasible) plus -10 non-positive pivot in the stage factorisation, -11 NaN iterates. */
int plm_qp_solve(plm_handle* h, int32_t batch, double* d_dx, int32_t* d_iters, int32_t* d_status, void* stream);
/* Read / write the persistent scaled iterates: x [batch][n], z [batch][m], y [batch][m]. */
int plm_qp_get_iterates(plm_handle* h, int32_t batch, double* d_x, double* d_z, double* d_y, void* stream);
int plm_qp_set_iterates(plm_handle* h, int32_t batch, const double* d_x, const double* d_z, const double* d_y, void* stream);
/* Scaling of the last plm_qp_update: D [batch][n], E [batch][m], c [batch] (diagnostics / tests). */
int plm_qp_get_scaling(plm_handle* h, int32_t batch, double* d_D, double* d_E, double* d_c, void* stream);

/* ---- Armijo line search (optimization/ocp.py:430-480) --------------------------------------------- */
/* d_info: [batch][4] = accepted (0/1), accepted step size, trials used, constraint-violation metric of the result */
int plm_line_search(plm_handle* h, const double* d_x, const double* d_p, const double* d_dx, int32_t batch,
                    double* d_x_new, double* d_info, void* stream);

/* ---- One full SQP iteration: body of the loop at optimization/ocp.py:383-406 ----------------------- */
/* d_stats: [batch][8] = qp iterations, qp status, accepted, step size, trials, f, g_metric, max violation */
int plm_sqp_step(plm_handle* h, const double* d_x, const double* d_p, int32_t batch, double* d_x_new,
                 double* d_stats, void* stream);

/* ---- One receding-horizon step, device resident: body of the generic loop of run_mpc.py:127-143 ------- */
/* update_gait_sequence(t) with t = t0[b] + t_add (d_t0 may be NULL: t0 = 0; utils/gait_sequence.py:37-77 and
 * optimization/ocp.py:234-242; gait 0 trot, 1 walk, 2 stand; dts_host = the `nodes` horizon step sizes of
 * optimization/ocp.py:71-74, host memory), then the starting point of the solve: x_mode 1 = warm_start() (forces in d_x
 * reset to the contact-masked f_des, the rest of the previous solution kept: ocp_whole_body_rnea.py:207-235), x_mode 0 =
 * no warm start (d_x <- opti.initial(): DX = 0, U = u_des, what the reference's solve() starts from when warm_start() is
 * not called, run_mpc.py:131-132), x_mode 2 = d_x as passed (first step); one SQP iteration d_x -> d_x_new (plm_sqp_step), and x_init <- integrate(x_init, DX_1 of d_x_new) inside d_p
 * (run_mpc.py:141); with update_tau_prev != 0 (whole_body_rnea, tau_nodes > 1) also tau_prev <- tau of node 1, as the
 * compiled-solver branch of the loop does (run_mpc.py:108-111; the generic branch keeps tau_prev).  d_x and d_p are
 * updated in place; the caller swaps d_x / d_x_new between steps. */
int plm_mpc_step(plm_handle* h, double* d_x, double* d_p, const double* d_t0, double t_add, int32_t gait, double gait_period,
                 const double* dts_host, int32_t x_mode, int32_t update_tau_prev, int32_t batch,
                 double* d_x_new, double* d_stats, void* stream);

/* Per-phase device times (ms) of the last plm_sqp_step on this handle: eval, qp_update, qp_solve, line_search.
 * Synchronises the stream. */
int plm_last_phase_ms(plm_handle* h, double* ms4);
/* Number of kernels this library has launched through the handle since creation. */
int64_t plm_launch_count(const plm_handle* h);
/* Measurement aid (the FP64 roofline denominator is measured on the box, not assumed): runs a DFMA
 * microbenchmark (eight independent fused multiply-add chains per thread, all SMs) on the current device and stores the
 * sustained FP64 rate in TFLOP/s (2 flops per DFMA).  Synchronises the device.  No reference counterpart. */
int plm_fp64_peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif
