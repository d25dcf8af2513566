#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample / shared-wavefront shares of one kernel of an .ncu-rep.

ncu's CSV source page is SASS only; the line of each SASS instruction comes from `nvdisasm -gi` of the object file the
kernel was built from (same build!).  usage: ncu_lines.py <ncu-rep> <kernel regex> <object file> <mangled name prefix> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", sys.argv[6] if len(sys.argv) > 6 else "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + mangled))
off2line, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
ia, ii, isamp, iw = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
base = None
agg, samp, wf = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iw or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    if a < base:      # next kernel instance
        break
    ln = off2line.get(a - base)
    agg[ln] += int(r[ii]); samp[ln] += int(r[isamp]); wf[ln] += int(r[iw] or 0)
tot, ts, tw = sum(agg.values()), max(1, sum(samp.values())), max(1, sum(wf.values()))
print("total warp instructions", tot, "samples", ts, "shared wavefronts", tw)
srcs = {}
order = samp.most_common(top) if os.environ.get("NCU_LINES_BY_SAMPLES") else agg.most_common(top)
for ln, _ in order:
    c = agg[ln]
    text = ""
    if ln:
        for d in (os.path.dirname(os.path.abspath(obj)),):
            f = os.path.join(d, ln[0])
            if os.path.exists(f):
                srcs.setdefault(f, open(f).read().splitlines())
                text = srcs[f][ln[1] - 1].strip()[:90]
    print(f"{100*c/tot:5.1f}% inst {100*samp[ln]/ts:5.1f}% samp {100*wf[ln]/tw:5.1f}% wf  {ln}  {text}")
