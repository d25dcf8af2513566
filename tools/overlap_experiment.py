#!/usr/bin/env python3
"""Does the QP update of one half of the batch hide behind the ADMM iterations of the other half?  Two handles of 4096
instances on two streams against one handle of 8192 (same instances, same arithmetic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp
from pino_locoman_b200.utils.robot import B2G

def make(batch):
    robot = B2G(); robot.set_gait_sequence("trot", 0.8)
    ocp = make_ocp(dynamics=bench.DYNAMICS, default_args=OCP_ARGS[bench.DYNAMICS], robot=robot, nodes=bench.NODES, solver="osqp", batch=batch, device="cuda:0")
    x, p = bench.synthetic_inputs(robot, ocp, batch, 0)
    ocp.init_solver()
    return ocp.handle, torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), ocp

B = 8192
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hf, xf, pf, _keep = make(B)
for _ in range(2):
    xf, _ = hf.sqp_step(xf, pf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    xf, _ = hf.sqp_step(xf, pf)
e1.record(); torch.cuda.synchronize()
print("one handle, 8192 instances: %.1f ms per step" % (e0.elapsed_time(e1) / 3))
del hf, xf, pf, _keep
torch.cuda.empty_cache()
parts = [make(B // nch) for _ in range(nch)]
streams = [torch.cuda.Stream() for _ in range(nch)]
xs = [p[1] for p in parts]
def step_all():
    for k, (h, _, p, _) in enumerate(parts):
        with torch.cuda.stream(streams[k]):
            xs[k], _ = h.sqp_step(xs[k], p)
for _ in range(2):
    step_all()
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    step_all()
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record(); torch.cuda.synchronize()
print("%d handles of %d instances on %d streams: %.1f ms per step of all" % (nch, B // nch, nch, e0.elapsed_time(e1) / 3))
