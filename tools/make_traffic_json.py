#!/usr/bin/env python3
"""profiles/traffic.json from ncu --set full captures: DRAM bytes and FP64-pipe fraction per kernel, per instance.

usage: make_traffic_json.py <rep>:<instances> [<rep>:<instances> ...] > profiles/traffic.json
Later reports override earlier ones for the same kernel; kernels whose counters came back NaN are skipped (the ADMM
kernel at 8192 instances overflows ncu's replay: it is taken from the 592-instance capture)."""
import csv, json, math, subprocess, sys
per, pipe, src, best_ms = {}, {}, {}, {}
for arg in sys.argv[1:]:
    rep, inst = arg.rsplit(":", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        try:
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            fp = float(r[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])
            ms = float(r[idx["gpu__time_duration.sum"]])
        except ValueError:
            continue
        if math.isnan(rd) or math.isnan(fp):
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        b = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
        if name == "qp_scale_kernel" and ms < 8.0 * int(inst) / 8192:      # the setup-time call (dummy data): not the update
            continue
        if name == "node_eval_kernel" and name in per and ms < best_ms.get(name, 0.0):      # residual-only line-search launches of the same kernel
            continue
        best_ms[name] = ms
        per[name], pipe[name], src[name] = b / int(inst), fp / 100.0, f"{rep.split('/')[-1]} ({inst} instances)"
print(json.dumps({"note": "DRAM traffic per instance (dram__bytes_read.sum + dram__bytes_write.sum) and FP64 pipe fraction "
                          "(sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active / 100) from ncu --set full captures; made by "
                          "tools/make_traffic_json.py; bench.py multiplies the bytes by the batch",
                  "source": src, "per_instance_bytes": per, "fp64_pipe_frac_ncu": pipe}, indent=1))
