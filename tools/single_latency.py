#!/usr/bin/env python3
"""Latency of one SQP iteration of a single B2 whole_body_rnea instance (BASELINE configs[1]): device time between
events, host wall time, and the per-phase event times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp
from pino_locoman_b200.utils.robot import B2
r1 = B2(); r1.set_gait_sequence("trot", 0.8)
o1 = make_ocp(dynamics="whole_body_rnea", default_args=OCP_ARGS["whole_body_rnea"], robot=r1, nodes=20, solver="osqp", batch=1, device="cuda:0")
o1.set_time_params(0.01, 0.08); o1.set_swing_params(0.07, [0.1, -0.2]); o1.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]))
o1.update_previous_torques(np.zeros(r1.nj)); o1.update_initial_state(o1.x_nom); o1.update_gait_sequence(0.0); o1.init_solver()
h = o1.handle
x1 = torch.from_numpy(o1.initial_guess()).cuda(); p1 = o1._p_device()
xn = torch.empty_like(x1); st = torch.empty(1, 8, dtype=torch.float64, device="cuda")
for k in range(10):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a0.record(); h.sqp_step(x1, p1, xn, st); a1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"call {k}: events {a0.elapsed_time(a1):.3f} ms  host enqueue {1e3*(t1-t0):.3f} ms  host total {1e3*(t2-t0):.3f} ms  phases {[round(v, 3) for v in h.last_phase_ms()]}  iters {st[0,0].item()}")
    x1, xn = xn, x1
