#!/usr/bin/env python3
"""In-kernel phase timing of the ADMM kernel (library built with -DPLM_ADMM_PROFILE): cycles of CTA 0."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pino_locoman_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "pino_locoman_b200", "libpinolocoman_b200_prof.so")
import torch, bench
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp
from pino_locoman_b200.utils.robot import B2G
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
DYN = sys.argv[2] if len(sys.argv) > 2 else bench.DYNAMICS
robot = B2G(); robot.set_gait_sequence("trot", 0.8)
ocp = make_ocp(dynamics=DYN, default_args=OCP_ARGS[DYN], robot=robot, nodes=bench.NODES, solver="osqp", batch=B, device="cuda:0")
x_host, p_host = bench.synthetic_inputs(robot, ocp, B, 0)
ocp.init_solver(); h = ocp.handle
x = torch.from_numpy(x_host).cuda(); p = torch.from_numpy(p_host).cuda()
x, stats = h.sqp_step(x, p); torch.cuda.synchronize()
lib = h.lib
out = (ctypes.c_longlong * 16)()
lib.plm_debug_admm_profile(out, 1)
x, stats = h.sqp_step(x, p); torch.cuda.synchronize()
lib.plm_debug_admm_profile(out, 0)
names = ["rhs(w,spmv_cols,add)", "fwd coupling", "panel wait + sched decode", "release (syncwarp/fence/atomic/refill)", "combine + barrier", "rect_panel loops (bwd)", "spmv_rows", "update", "check", "sym_panel loops (fwd)", "cpart store + barrier", "coupling: own work (barrier wait is in fwd coupling)", "", "", "", "loop"]
tot = sum(out)
print("phase ms", h.last_phase_ms(), "iters", stats[0, 0].item())
for n_, v in zip(names, out):
    if v: print(f"{n_:24s} {v:12d} cycles  {100.0 * v / tot:5.1f}%   per-iter {v / max(1.0, stats[0,0].item()):9.0f}")
