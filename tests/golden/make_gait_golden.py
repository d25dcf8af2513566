#!/usr/bin/env python3
"""Generate gait-schedule / swing-spline golden vectors by running the REFERENCE's own
utils/gait_sequence.py (importable here with a stub casadi module).

Run in the build container only (needs /root/reference):
    python tests/golden/make_gait_golden.py
Floats are stored as hex strings so the comparison is bit-exact.
"""
import importlib.util
import json
import os
import sys
import types

REF = "/root/reference/utils/gait_sequence.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gait_golden.json")


def load_reference():
    ca = types.ModuleType("casadi")
    ca.if_else = lambda c, a, b: a if c else b   # scalar evaluation of the symbolic if_else
    sys.modules["casadi"] = ca
    spec = importlib.util.spec_from_file_location("ref_gait_sequence", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    gs = load_reference()
    cases = []
    for gait, period, nodes, dt_min, dt_max, ks in [
        ("trot", 0.8, 20, 0.01, 0.08, list(range(0, 80)) + [123, 400, 7919]),
        ("trot", 0.8, 14, 0.01, 0.08, list(range(0, 80, 3))),
        ("trot", 0.5, 20, 0.01, 0.05, list(range(0, 50, 7))),
        ("walk", 0.8, 20, 0.01, 0.08, list(range(0, 80))),
        ("walk", 1.2, 14, 0.02, 0.08, list(range(0, 60, 5))),
        ("stand", 0.8, 20, 0.01, 0.08, [0, 17]),
    ]:
        gamma = (dt_max / dt_min) ** (1 / (nodes - 1))       # optimization/ocp.py:72-74
        dts = [dt_min * gamma ** i for i in range(nodes)]
        seq = gs.GaitSequence(gait, period)
        for k in ks:
            contact, swing = seq.get_gait_schedule(k * dt_min, dts, nodes)
            cases.append(dict(gait=gait, period=period, nodes=nodes, dt_min=dt_min, dt_max=dt_max, k=k,
                              n_contacts=seq.n_contacts, swing_period=float(seq.swing_period).hex(),
                              dts=[float(d).hex() for d in dts],
                              contact=[[int(c) for c in row] for row in contact],
                              swing=[[float(s).hex() for s in row] for row in swing]))
    splines = []
    for phase in [0.0, 0.025, 0.1, 0.25, 0.3, 0.49999, 0.5, 0.50001, 0.75, 0.9, 1.0]:
        for (T, h, vl, vt) in [(0.4, 0.07, 0.1, -0.2), (0.2, 0.1, 0.1, -0.2), (0.8, 0.05, 0.0, -0.1)]:
            v = gs.get_spline_vel_z(phase, T, h, vl, vt)
            splines.append(dict(phase=phase, swing_period=T, h_max=h, v_liftoff=vl, v_touchdown=vt,
                                vel_z=float(v).hex()))
    with open(OUT, "w") as f:
        json.dump(dict(source="lukasmolnar/pino-locoman utils/gait_sequence.py (GaitSequence, get_spline_vel_z)",
                       schedules=cases, splines=splines), f)
    print("wrote", OUT, len(cases), "schedules", len(splines), "spline samples")


if __name__ == "__main__":
    main()
