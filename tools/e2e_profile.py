import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bench
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp
from pino_locoman_b200.utils.robot import B2G
B = 8192
robot = B2G(); robot.set_gait_sequence("trot", 0.8)
ocp = make_ocp(dynamics=bench.DYNAMICS, default_args=OCP_ARGS[bench.DYNAMICS], robot=robot, nodes=bench.NODES, solver="osqp", batch=B, device="cuda:0")
x_host, p_host = bench.synthetic_inputs(robot, ocp, B, 0)
ocp.init_solver()
ocp._x0 = x_host
for _ in range(2): ocp.solve(retract_all=False)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(3): ocp.solve(retract_all=False)
torch.cuda.synchronize()
print("per solve ms", (time.perf_counter() - t0) / 3 * 1e3, "device ms", sum(ocp.handle.last_phase_ms()))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
