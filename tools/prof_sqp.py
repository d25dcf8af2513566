#!/usr/bin/env python3
"""Small driver for profiling: a few SQP steps of the bench workload at a reduced batch."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from pino_locoman_b200 import OCP_ARGS  # noqa: E402
from pino_locoman_b200.optimization import make_ocp  # noqa: E402
from pino_locoman_b200.utils.robot import B2G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--dynamics", default=bench.DYNAMICS)
args = ap.parse_args()
robot = B2G()
robot.set_gait_sequence("trot", 0.8)
ocp = make_ocp(dynamics=args.dynamics, default_args=OCP_ARGS[args.dynamics], robot=robot, nodes=bench.NODES, solver="osqp",
               batch=args.batch, device="cuda:0")
x_host, p_host = bench.synthetic_inputs(robot, ocp, args.batch, 0)
ocp.init_solver()
h = ocp.handle
x = torch.from_numpy(x_host).cuda()
p = torch.from_numpy(p_host).cuda()
for _ in range(args.steps):
    x, stats = h.sqp_step(x, p)
torch.cuda.synchronize()
print("phase ms", h.last_phase_ms(), "qp iters", stats[:, 0].mean().item())
