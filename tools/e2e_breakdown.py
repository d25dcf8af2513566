#!/usr/bin/env python3
"""Where the end-to-end step (OCP.solve with host buffers) spends its time beyond the device work."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp
from pino_locoman_b200.utils.robot import B2G
B = 8192
robot = B2G(); robot.set_gait_sequence("trot", 0.8)
ocp = make_ocp(dynamics=bench.DYNAMICS, default_args=OCP_ARGS[bench.DYNAMICS], robot=robot, nodes=bench.NODES, solver="osqp", batch=B, device="cuda:0")
x, p = bench.synthetic_inputs(robot, ocp, B, 0)
ocp.init_solver()
ocp.set_initial(x)
for _ in range(2):
    ocp.solve(retract_all=False)
h = ocp.handle
cur = ocp._x0
for rep in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    ocp._pin_x.numpy()[:] = cur; t.append(time.perf_counter())
    xd = ocp._pin_x.to(h.device, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    pd = ocp._p_device(); torch.cuda.synchronize(); t.append(time.perf_counter())
    xn, st = h.sqp_step(xd, pd); torch.cuda.synchronize(); t.append(time.perf_counter())
    out = ocp._pin_out[0]; out.copy_(xn, non_blocking=True); ocp._pin_stats.copy_(st, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    ocp.retract_stacked_sol(out.numpy(), False); t.append(time.perf_counter())
    names = ["host copy into pinned", "H2D x", "p (cached)", "sqp_step", "D2H", "retract_stacked_sol"]
    print("  ".join(f"{n} {1e3*(b-a):.1f} ms" for n, a, b in zip(names, t[:-1], t[1:])))
t0 = time.perf_counter()
torch.from_numpy(ocp._pin_x.numpy()).copy_(torch.from_numpy(cur)); print("torch host copy %.1f ms" % (1e3 * (time.perf_counter() - t0)), "threads", torch.get_num_threads())
