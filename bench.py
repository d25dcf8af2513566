#!/usr/bin/env python3
"""Benchmark of the hot path: batched SQP iterations (dyn+Jacobian node evaluation -> ADMM QP -> Armijo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[4] per-GPU share, weak scaling): B2G (B2 + Z1 arm) whole_body_rnea, trot, N=20
nodes, `batch` independent MPC instances per GPU with the synthetic state distributions of SURVEY.md 8(d)
(seed 1234 + 1000*rank).  One "step" = one full SQP iteration of every instance (optimization/ocp.py:383-406);
consecutive steps continue the SQP sequence (x <- x_new, warm-started ADMM iterates), as the MPC loop does.

Prints ONE JSON line on rank 0.  `value` = SQP iterations/s with inputs resident in HBM; `e2e` = the same through
the plugin surface (OCP.solve with host buffers, H2D + D2H inside the timed region); `node_evals_per_s` and the two
roofline objects report the dyn+Jacobian kernel and the dominant (ADMM) kernel against the measured HBM peak;
`cpu_baseline` is the compiled C++ port of the reference algorithm (oracle/cport) on one host core, with the numpy oracle
beside it (ports, for context only: the real casadi / pinocchio / OSQP stack is not installable here).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ROBOT, DYNAMICS, NODES = "b2g", "whole_body_rnea", 20
# ALGORITHMIC bytes per node-eval / residual-only eval, B2G whole_body_rnea, avg over N=20
BYTES_NODE_EVAL = 10464.0
BYTES_NODE_RESID = 2220.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def synthetic_inputs(robot, ocp, batch, rank):
    """x [B,n], p [B,np] on the host, distributions of SURVEY.md 8(d)."""
    rng = np.random.default_rng(1234 + 1000 * rank)
    B, nj = batch, robot.nj
    q = np.tile(robot.q0, (B, 1))
    q[:, 0:2] = rng.uniform(-1, 1, (B, 2))
    q[:, 2] = rng.uniform(0.28, 0.40, B) if robot.q0[2] < 0.4 else rng.uniform(0.45, 0.65, B)      # Go2 stands lower
    yaw, roll, pitch = rng.uniform(-np.pi, np.pi, B), rng.uniform(-0.3, 0.3, B), rng.uniform(-0.3, 0.3, B)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    q[:, 3] = sr * cp * cy - cr * sp * sy
    q[:, 4] = cr * sp * cy + sr * cp * sy
    q[:, 5] = cr * cp * sy - sr * sp * cy
    q[:, 6] = cr * cp * cy + sr * sp * sy
    qj = rng.uniform(robot.joint_pos_min, robot.joint_pos_max, (B, nj))
    q[:, 7:] = robot.q0[7:] + 0.9 * (qj - robot.q0[7:])
    v = np.concatenate((rng.uniform(-1, 1, (B, 6)), rng.uniform(-0.25, 0.25, (B, nj)) * robot.joint_vel_max), 1)
    ocp.set_time_params(0.01, 0.08)
    ocp.set_swing_params(0.07, [0.1, -0.2])
    ocp.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), rng.uniform(-20, 20, (B, 3)), rng.uniform(-0.2, 0.2, (B, 3)))
    if ocp.dynamics == "centroidal_vel":      # state (h, q): normalised centroidal momentum instead of v
        ocp.update_initial_state(np.concatenate((rng.uniform(-0.5, 0.5, (B, 6)), q), 1))
    else:
        ocp.update_initial_state(np.concatenate((q, v), 1))
    ocp.update_gait_sequence(rng.integers(0, 80, B) * 0.01)
    if hasattr(ocp, "update_previous_torques"):
        ocp.update_previous_torques(np.zeros(nj))
    x = ocp.initial_guess()
    h = ocp.handle
    contact = ocp._get("contact_schedule").reshape(B, ocp.nodes, 4)
    mg = ocp.mass * 9.81
    lead = ocp._lead()                         # leading input block: a (nv), v (nv) or tau_j (nj)
    for i in range(ocp.nodes + 1):
        o = h.x_off[i]
        if i > 0:
            x[:, o:o + ocp.ndx_opt] = rng.normal(0, 0.05, (B, ocp.ndx_opt))
        if i == ocp.nodes:
            break
        u = x[:, o + ocp.ndx_opt:o + ocp.ndx_opt + ocp.nu_opt[i]]
        if ocp.dynamics == "whole_body_aba":
            u[:, :lead] = rng.uniform(-0.5, 0.5, (B, lead)) * robot.joint_torque_max
        elif ocp.dynamics == "centroidal_vel":
            u[:, :lead] = v
        else:
            u[:, :lead] = rng.normal(0, 5, (B, lead))
        for k in range(4):
            fz = rng.uniform(0, mg, B)
            u[:, lead + 3 * k] = 0.7 * fz * rng.uniform(-0.5, 0.5, B) * contact[:, i, k]
            u[:, lead + 3 * k + 1] = 0.7 * fz * rng.uniform(-0.5, 0.5, B) * contact[:, i, k]
            u[:, lead + 3 * k + 2] = fz * contact[:, i, k]
        if robot.nf > 12:
            u[:, lead + 12:lead + 15] = rng.uniform(-20, 20, (B, 3))
        if ocp.nu_opt[i] > lead + robot.nf:
            u[:, lead + robot.nf:] = rng.uniform(-0.5, 0.5, (B, nj)) * robot.joint_torque_max
    return x, ocp._p.copy()


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the numpy oracle (a port of the reference algorithm; see oracle/__init__.py)
# ---------------------------------------------------------------------------------------------------------------
def _oracle_sqp_worker(args):
    x, p, iters = args
    from oracle.model import OracleRobot
    from oracle.ocp import OracleOCP
    from oracle.sqp import OracleSQP
    o = OracleOCP(OracleRobot(ROBOT), DYNAMICS, NODES)
    for name, (off, sz) in o.p_layout.items():
        o.params[name][:] = p[off:off + sz]
    s = OracleSQP(o)
    s.init_solver()
    t_eval = t_all = 0.0
    for _ in range(iters):
        t0 = time.perf_counter()
        o.sqp_data(x, p)
        t_eval += time.perf_counter() - t0
        t0 = time.perf_counter()
        x, _ = s.solve(x, p)
        t_all += time.perf_counter() - t0
    return t_eval, t_all


def cpu_oracle_timing(x, p, n_inst, iters, procs):
    jobs = [(x[i], p[i], iters) for i in range(n_inst)]
    t0 = time.perf_counter()
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_oracle_sqp_worker, jobs)
    else:
        res = [_oracle_sqp_worker(j) for j in jobs]
    wall = time.perf_counter() - t0
    t_eval = sum(r[0] for r in res)
    t_all = sum(r[1] for r in res)
    return wall, t_eval, t_all


def oracle_synthetic_inputs(n_inst, seed):
    """Synthetic instances of the bench workload built with the oracle only (SURVEY 8(d) distributions): the reference arm
    must not map the product library."""
    from emu_util import random_problem
    from oracle.model import OracleRobot
    from oracle.ocp import OracleOCP
    rng = np.random.default_rng(seed)
    o = OracleOCP(OracleRobot(ROBOT), DYNAMICS, NODES)
    xs, ps = zip(*[random_problem(o, rng) for _ in range(n_inst)])
    return np.stack(xs), np.stack(ps)


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm on all host cores.  casadi / pinocchio / OSQP are not installable
    here or on the GPU box (profiles/probe_real_stack_r02.log), so what is timed is the restated algorithm: the compiled
    C++ port (oracle/cport, -O3 -march=native, one process per core) when it is built, else the numpy oracle."""
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    n_inst = 8 * procs            # bounded sample: eight instances per host core per step (~0.3 s of CPU work per core)
    x, p = oracle_synthetic_inputs(n_inst, 1234)
    cport = _load_cport()
    walls = []
    for step in range(args.warmup + args.steps):
        if cport is not None:
            t_all = cport_timing(x, p, n_inst, procs)[0]
        else:
            t_all = cpu_oracle_timing(x, p, n_inst, 1, procs)[2] / procs
        if step >= args.warmup:
            walls.append(t_all)      # wall seconds for n_inst SQP iterations on `procs` cores
    ms = 1e3 * float(np.mean(walls))
    value = n_inst / (ms / 1e3)
    kind = "port (C++)" if cport is not None else "port"
    line = {"impl": "reference", "metric": "sqp_iters_per_s", "value": value, "unit": "SQP iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{ROBOT} {DYNAMICS} trot N={NODES}, SQP iteration (sqp_data + OSQP + Armijo)", "robot": ROBOT,
                       "dynamics": DYNAMICS, "nodes": NODES, "instances_per_step": n_inst,
                       "same_config": False,
                       "note": "bounded sample: eight instances per host core per step (the GPU arm runs 8192 per GPU); a restated "
                               "port, not the real casadi/pinocchio/OSQP stack"},
            "cpu_baseline": {"value": value, "unit": "SQP iters/s", "cores": procs, "kind": kind,
                             "sample": f"{n_inst} instances x 1 SQP iteration per step, one process per core, "
                                       + ("C++ port of the reference algorithm (-O3 -march=x86-64-v3)" if cport is not None else "numpy oracle")
                                       + " (restated reference algorithm; casadi/pinocchio/osqp not installable)"},
            "e2e": {"value": value, "unit": "SQP iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _load_cport():
    """The compiled CPU port (oracle/cport), or None when it has not been built."""
    try:
        from oracle.cport import load
        return load()
    except Exception:
        return None


def cport_timing(x, p, n_inst, procs):
    """n_inst SQP iterations (one per instance) with the C++ port on `procs` processes; (wall seconds, node-eval seconds)."""
    from oracle.cport import time_sqp
    return time_sqp(ROBOT, DYNAMICS, NODES, x, p, procs)


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8192, help="MPC instances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra legs on BASELINE configs[0], [2], [3]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.handle import _ptr
    from pino_locoman_b200.optimization import make_ocp
    from pino_locoman_b200.sharding import gather_instance_results
    from pino_locoman_b200.utils.robot import B2G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pino_locoman_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, W, K = args.batch, max(args.warmup, 0), args.steps

    robot = B2G()
    robot.set_gait_sequence("trot", 0.8)
    ocp = make_ocp(dynamics=DYNAMICS, default_args=OCP_ARGS[DYNAMICS], robot=robot, nodes=NODES, solver="osqp", batch=B, device=dev)
    h = ocp.handle
    x_host, p_host = synthetic_inputs(robot, ocp, B, rank)
    ocp.init_solver()
    x = torch.from_numpy(x_host).to(dev)
    p = torch.from_numpy(p_host).to(dev)
    x_new = torch.empty_like(x)
    stats = torch.empty(B, 8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        nonlocal x, x_new
        h.sqp_step(x, p, x_new, stats)
        if world > 1:   # the only exchange of the path: per-instance costs to every rank
            gather_instance_results(stats[:, 5].contiguous(), B * world)
        x, x_new = x_new, x

    # ---- device-resident timing
    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    phase = np.zeros(4)
    qp_iters, trials, accepted = [], [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step()
        phase += np.array(h.last_phase_ms())     # synchronises on the step's last event
        s = stats.cpu().numpy()
        qp_iters.append(float(s[:, 0].mean()))
        trials.append(float(s[:, 4].mean()))
        accepted.append(float(s[:, 2].mean()))
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = h.launch_count() - launches0
    launches_timed = launches          # kernels of this library launched inside the timed region (K steps)
    # ---- dyn+Jacobian node-eval kernel alone (sqp_data without objective / bounds)
    g = torch.empty(B, h.m, dtype=torch.float64, device=dev)
    J = torch.empty(B, h.nnz, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3):
        h.lib.plm_sqp_data(h._h, _ptr(x), _ptr(p), B, None, _ptr(J), _ptr(g), None, None, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        h.lib.plm_sqp_data(h._h, _ptr(x), _ptr(p), B, None, _ptr(J), _ptr(g), None, None, stream)
    e1.record()
    torch.cuda.synchronize()
    ms_eval = e0.elapsed_time(e1) / reps
    launches += reps
    del g, J
    # ---- north-star operating point: 512 instances per GPU (4096 on 8 GPUs), one full SQP iteration, latency
    Bt = min(512, B)
    xt, pt = x[:Bt].clone(), p[:Bt].contiguous()
    xt_new, st_t = torch.empty_like(xt), torch.empty(Bt, 8, dtype=torch.float64, device=dev)
    t_ms, t_phase, t_iters = [], [], []
    for k in range(4):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        h.sqp_step(xt, pt, xt_new, st_t)
        a1.record()
        torch.cuda.synchronize()
        t_ms.append(a0.elapsed_time(a1))
        t_phase.append(h.last_phase_ms())
        t_iters.append(float(st_t[:, 0].mean()))
        xt, xt_new = xt_new, xt
    target_ms = float(np.median(t_ms[1:]))
    target_phase = [float(v) for v in np.median(np.array(t_phase[1:]), axis=0)]
    target_iters = float(np.mean(t_iters[1:]))
    launches += h.launch_count() - launches0 - launches
    del xt, pt, xt_new, st_t
    # ---- closed loop: device-resident receding-horizon steps of every instance (gait update, warm
    # start, one SQP iteration, state advance), nominal start states, extra key
    from pino_locoman_b200.mpc import BatchedMPC
    p_saved = ocp._p.copy()                 # the end-to-end leg below runs on the synthetic parameters again
    ocp.update_initial_state(ocp.x_nom)
    ocp._x0 = None
    mpc = BatchedMPC(ocp, warm_start=True, t0=np.arange(B) % 80 * 0.01)
    for _ in range(2):
        mpc.step()
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mpc_steps = 3
    m0.record()
    for _ in range(mpc_steps):
        mpc.step()
    m1.record()
    torch.cuda.synchronize()
    ms_mpc = m0.elapsed_time(m1) / mpc_steps
    mpc_iters = float(mpc.stats[:, 0].mean())
    launches += h.launch_count() - launches0 - launches
    del mpc
    ocp._p[:] = p_saved
    ocp._p_dirty = True
    # ---- BASELINE configs[1]: single B2 whole_body_rnea instance, latency of one SQP iteration (parity-test case, extra key)
    single_ms = None
    if rank == 0:
        from pino_locoman_b200.utils.robot import B2
        r1 = B2()
        r1.set_gait_sequence("trot", 0.8)
        o1 = make_ocp(dynamics=DYNAMICS, default_args=OCP_ARGS[DYNAMICS], robot=r1, nodes=NODES, solver="osqp", batch=1, device=dev)
        o1.set_time_params(0.01, 0.08)
        o1.set_swing_params(0.07, [0.1, -0.2])
        o1.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]))
        o1.update_previous_torques(np.zeros(r1.nj))
        o1.update_initial_state(o1.x_nom)
        o1.update_gait_sequence(0.0)
        o1.init_solver()
        x1 = torch.from_numpy(o1.initial_guess()).to(dev)
        p1 = o1._p_device()
        ts, phs, its1 = [], [], []
        for k in range(7):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            x1, st1 = o1.handle.sqp_step(x1, p1)
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
            phs.append([float(v) for v in o1.handle.last_phase_ms()])
            its1.append(float(st1[0, 0]))
        kmed = 1 + int(np.argsort(ts[1:])[len(ts[1:]) // 2])       # the call with the median time (ADMM iterations vary: 25 .. 100)
        single_ms, single_phase, single_iters = float(ts[kmed]), phs[kmed], its1[kmed]
        launches += o1.handle.launch_count()
        del o1
    # ---- the other BASELINE configs (parity-test cases; measured here as extra keys, rank 0 only):
    # [0] Go2 centroidal_vel single instance, [2] B2G whole_body_aba x 1024 instances, [3] 4096 B2 centroidal_acc node sweep
    other = {}
    if rank == 0 and not args.no_other_configs:
        from pino_locoman_b200.utils.robot import B2, Go2

        def config_leg(robot_cls, dynamics, batch, sweep_only, bytes_node_eval, gait="trot"):
            nonlocal launches
            r = robot_cls()
            r.set_gait_sequence(gait, 0.8)
            o = make_ocp(dynamics=dynamics, default_args=OCP_ARGS[dynamics], robot=r, nodes=NODES, solver="osqp", batch=batch, device=dev)
            xh, ph = synthetic_inputs(r, o, batch, 0)
            o.init_solver()
            hh = o.handle
            xd, pd = torch.from_numpy(xh).to(dev), torch.from_numpy(ph).to(dev)
            gg = torch.empty(batch, hh.m, dtype=torch.float64, device=dev)
            JJ = torch.empty(batch, hh.nnz, dtype=torch.float64, device=dev)
            for _ in range(3):
                hh.lib.plm_sqp_data(hh._h, _ptr(xd), _ptr(pd), batch, None, _ptr(JJ), _ptr(gg), None, None, stream)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                hh.lib.plm_sqp_data(hh._h, _ptr(xd), _ptr(pd), batch, None, _ptr(JJ), _ptr(gg), None, None, stream)
            c1.record()
            torch.cuda.synchronize()
            ms_sw = c0.elapsed_time(c1) / 10
            res = {"instances": batch, "n": hh.n, "m": hh.m, "nnz_J": hh.nnz, "ms_per_sweep": ms_sw,
                   "node_evals_per_s": batch * NODES / (ms_sw / 1e3), "bytes_per_node_eval": bytes_node_eval,
                   "roofline_frac_hbm": batch * NODES * bytes_node_eval / (ms_sw / 1e3) / 1e9 / peak_hbm}
            if not sweep_only:
                xn, st = torch.empty_like(xd), torch.empty(batch, 8, dtype=torch.float64, device=dev)
                ts, its = [], []
                for k in range(5):
                    c0.record()
                    hh.sqp_step(xd, pd, xn, st)
                    c1.record()
                    torch.cuda.synchronize()
                    ts.append(c0.elapsed_time(c1))
                    its.append(float(st[:, 0].mean()))
                    xd, xn = xn, xd
                ms_it = float(np.median(ts[1:]))
                res.update({"ms_per_sqp_iter": ms_it, "sqp_iters_per_s": batch / (ms_it / 1e3), "admm_iters_avg": float(np.mean(its[1:])),
                            "phase_ms": dict(zip(("eval", "qp_update", "qp_solve", "line_search"), [float(v) for v in hh.last_phase_ms()]))})
            launches += hh.launch_count()
            return res

        peak_hbm = load_peaks()[0]
        other["go2_centroidal_vel_1"] = config_leg(Go2, "centroidal_vel", 1, False, 6144.0)
        other["b2g_whole_body_aba_1024"] = config_leg(B2G, "whole_body_aba", 1024, False, 20172.0)
        other["b2_centroidal_acc_4096_sweep"] = config_leg(B2, "centroidal_acc", 4096, True, 7320.0)
        # the other gaits of utils/gait_sequence.py:53-75 on the bench formulation
        other["b2g_whole_body_rnea_1024_walk"] = config_leg(B2G, DYNAMICS, 1024, False, BYTES_NODE_EVAL, gait="walk")
        other["b2g_whole_body_rnea_1024_stand"] = config_leg(B2G, DYNAMICS, 1024, False, BYTES_NODE_EVAL, gait="stand")
    # ---- end to end through the plugin surface: host buffers in, host buffers out
    # every step hands over that step's inputs from the host, as the reference's loop does (run_mpc.py:127-133: x_init,
    # schedules -> parameters; starting point): p [B, np] and x [B, n] go host -> device, x_new and the statistics come back
    ocp.set_initial(x.cpu().numpy())
    x_init_host = ocp._get("x_init").copy()
    for _ in range(min(W, 1)):
        ocp.update_initial_state(x_init_host)
        ocp.solve(retract_all=False)
    barrier()
    t0 = time.perf_counter()
    ke = max(1, min(K, 3))
    for _ in range(ke):
        ocp.update_initial_state(x_init_host)      # marks the parameter image dirty: p is uploaded again
        ocp.solve(retract_all=False)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / ke
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- reduce over ranks (max time)
    tt = torch.tensor([ms_total, ms_eval, t_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, ms_eval, ms_e2e = [float(v) for v in tt.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / K
    total_inst = B * world
    value = total_inst / (ms_per_step / 1e3)
    peak, peak_src = load_peaks()
    Kavg, Tavg = float(np.mean(qp_iters)), float(np.mean(trials))
    nnzF = h.dims.kkt_factor_doubles
    n, m = h.n, h.m
    ms_admm = phase[2] / K
    bytes_admm = B * Kavg * (8.0 * nnzF + 8.0 * (3 * n + 4 * m))       # K (8 nnz(F) + 8 (3n + 4m)) per instance
    ach_admm = bytes_admm / (ms_admm / 1e3) / 1e9
    # the same with the minimal factor (packed S_i^-1 only; the stored factor also holds the back-substitution blocks B_i)
    nnzF_min = sum((h.x_off[i + 1] - h.x_off[i]) * (h.x_off[i + 1] - h.x_off[i] + 1) // 2 for i in range(NODES)) + h.ndx * (h.ndx + 1) // 2
    ach_admm_min = B * Kavg * (8.0 * nnzF_min + 8.0 * (3 * n + 4 * m)) / (ms_admm / 1e3) / 1e9
    ach_eval = B * NODES * BYTES_NODE_EVAL / (ms_eval / 1e3) / 1e9
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):     # per-instance DRAM bytes from the committed ncu captures, scaled to this batch
        with open(tpath) as f:
            traffic = {k: v * B for k, v in json.load(f).get("per_instance_bytes", {}).items()}
    fp64_peak = ctypes.c_double(0.0)       # DFMA microbenchmark on this GPU
    h.lib.plm_fp64_peak(ctypes.byref(fp64_peak))
    ncu_pipe = {}
    if os.path.exists(tpath):
        with open(tpath) as f:
            ncu_pipe = json.load(f).get("fp64_pipe_frac_ncu", {})
    line = {
        "metric": "sqp_iters_per_s", "value": value, "unit": "SQP iters/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{ROBOT} {DYNAMICS} trot N={NODES}, {B} MPC instances per GPU, one SQP iteration per step "
                               "(BASELINE configs[4] per-GPU share; configs[1] single-instance B2 is a parity-test case)",
                   "robot": ROBOT, "dynamics": DYNAMICS, "nodes": NODES, "instances_per_gpu": B, "n": n, "m": m, "nnz_J": h.nnz,
                   "l2": "per-step inputs (J, scaled A, QP factor: > 10 GB per GPU) exceed the 126 MB L2",
                   "collective": "all_gather of per-instance cost" if world > 1 else "none"},
        "node_evals_per_s": total_inst * NODES / (ms_eval / 1e3),
        "phase_ms": {"eval": phase[0] / K, "qp_update": phase[1] / K, "qp_solve": phase[2] / K, "line_search": phase[3] / K},
        "qp": {"admm_iters_avg": Kavg, "line_search_trials_avg": Tavg, "accepted_frac": float(np.mean(accepted)), "nnz_F": nnzF},
        "roofline": {"kernel": "qp_admm_kernel", "bound": "hbm", "achieved": ach_admm, "peak": peak, "unit": "GB/s",
                     "frac": ach_admm / peak, "traffic": traffic.get("qp_admm_kernel"),
                     "traffic_source": "profiles/traffic.json: dram__bytes per instance of this round's kernels (ncu --set full, one wave of "
                                       "592 instances; the counters overflow at 8192) x the batch: not measured in this run",
                     "peak_source": peak_src,
                     "algorithmic_bytes": bytes_admm, "nnz_F_stored": nnzF, "nnz_F_minimal": nnzF_min,
                     "frac_minimal_factor": ach_admm_min / peak, "ms_per_step_kernel": ms_admm,
                     # the same time against the DRAM bytes ncu counts for the kernel (they include the two streams of the
                     # scaled constraint matrix, which SURVEY 8d's formula leaves out)
                     "frac_of_peak_with_ncu_traffic": (traffic["qp_admm_kernel"] / (ms_admm / 1e3) / 1e9 / peak
                                                       if traffic.get("qp_admm_kernel") else None)},
        "roofline_node_eval": {"kernel": "node_eval_kernel", "bound": "hbm", "achieved": ach_eval, "peak": peak, "unit": "GB/s",
                               "frac": ach_eval / peak, "traffic": traffic.get("node_eval_kernel"),
                               "bytes_per_node_eval": BYTES_NODE_EVAL, "ms_per_sweep": ms_eval,
                               "fp64_peak_tflops_measured": fp64_peak.value, "fp64_pipe_frac_ncu": ncu_pipe.get("node_eval_kernel")},
        "mpc": {"workload": "closed loop from the nominal state: gait update + warm start + 1 SQP iteration + state advance, device resident",
                "mpc_steps_per_s": B / (ms_mpc / 1e3), "ms_per_step": ms_mpc, "admm_iters_avg": mpc_iters},
        "e2e": {"value": total_inst / (ms_e2e / 1e3), "unit": "SQP iters/s", "h2d_bytes_per_step": int(B * (n + h.np) * 8),
                "d2h_bytes_per_step": int(B * (n + 8) * 8)},
        "target": {"workload": f"north-star operating point: {Bt} {ROBOT} {DYNAMICS} N={NODES} instances per GPU (4096 on 8 GPUs), one full SQP "
                               "iteration; the north-star asks for < 1 ms",
                   "instances_per_gpu": Bt, "ms_per_sqp_iter": target_ms,
                   "phase_ms": dict(zip(("eval", "qp_update", "qp_solve", "line_search"), target_phase)), "admm_iters_avg": target_iters,
                   "hbm_floor_ms": Bt * target_iters * (8.0 * nnzF + 8.0 * (3 * n + 4 * m)) / (peak * 1e9) * 1e3,
                   "hbm_floor_note": "instances x ADMM iterations x (8 nnz(F) + 8 (3n + 4m)) bytes / measured HBM peak: the factor of an "
                                     "instance (1.36 MB) is re-read every iteration and 512 of them (0.7 GB) do not fit the 126 MB L2"},
        "single_instance": {"workload": "b2 whole_body_rnea trot N=20, 1 instance (BASELINE configs[1])", "ms_per_sqp_iter": single_ms,
                            "phase_ms": dict(zip(("eval", "qp_update", "qp_solve", "line_search"), single_phase)), "admm_iters": single_iters},
        "other_configs": other,
        "gpu_launches": int(launches_timed), "gpu_launches_all_legs": int(launches),
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        # CPU baseline on ONE host core, bounded sample of the same workload (same synthetic instances as the GPU arm):
        # the compiled C++ port of the reference algorithm (oracle/cport) when it is built, plus the numpy oracle
        n_np = 4
        wall, t_eval, t_all = cpu_oracle_timing(x_host, p_host, n_np, 1, 1)
        numpy_port = {"value": n_np / t_all, "unit": "SQP iters/s", "cores": 1, "node_evals_per_s": n_np * NODES / t_eval,
                      "sample": f"{n_np} instances x 1 SQP iteration, numpy oracle"}
        if _load_cport() is not None:
            n_cpu = 256        # ~10 s of CPU work on one core
            t_sqp, t_ev = cport_timing(x_host[:n_cpu], p_host[:n_cpu], n_cpu, 1)
            line["cpu_baseline"] = {"value": n_cpu / t_sqp, "unit": "SQP iters/s", "cores": 1, "kind": "port (C++)",
                                    "node_eval_share": t_ev / t_sqp,
                                    "sample": f"{n_cpu} instances x 1 SQP iteration of the same workload, C++ port of the reference algorithm "
                                              "(-O3, one core; restated, not casadi/pinocchio/OSQP: those are not installable on this box, "
                                              "profiles/probe_real_stack_r02.log)",
                                    "numpy_port": numpy_port}
        else:
            line["cpu_baseline"] = dict(numpy_port, kind="port")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
