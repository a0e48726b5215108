#!/usr/bin/env python
"""Aggregate an ncu SASS source page per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv libswarmstep.so 'swarm_kernelILi3ELb1ELi24ELi0E' [top_n]

Joins instruction offsets with `nvdisasm -g` line info of the matching kernel in the .so and prints
executed warp-instructions and stall samples per source line (profiling aid, not product code).
"""
import csv, os, re, subprocess, sys, tempfile, collections

sass_csv, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
HELPER_MAX = int(os.environ.get("HELPER_MAX", "0"))  # attribute lines <= HELPER_MAX (tiny inlined helpers) to their call site
line_of = {}
cur_fn, cur_line, active = None, None, False
chain = []
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        active = pat in m.group(1) and "$" not in m.group(1)
        cur_line = None
        continue
    if not active:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        chain.append(int(m.group(2)) if m.group(1).endswith("swarm_step.cu") else None)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m:
        pick = None
        for c in chain:          # innermost first; skip helper-definition lines and other files
            if c is not None and c > HELPER_MAX:
                pick = c
                break
        if pick is None and chain:
            pick = chain[0]
        if chain:
            cur_line = pick
        chain = []
        line_of[int(m.group(1), 16)] = (cur_line, m.group(2))
rows = list(csv.reader(open(sass_csv)))
# first kernel block only
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_idx[0]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
hdr = rows[start]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
base = int(body[0][ia], 16)
per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_e = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    line, op = line_of.get(off, (None, "?"))
    e, s = int(r[ie] or 0), int(r[isamp] or 0)
    tot_e += e; tot_s += s
    pl = per_line[line]
    pl[0] += e; pl[1] += s
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            pl[2][hdr[i]] += v
src = open(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "swarm_step.cu")).read().splitlines()
print(f"total warp-instr {tot_e}  samples {tot_s}  sass instrs {len(body)}")
SORT = 0 if os.environ.get("SORT", "samples") == "inst" else 1
for line, (e, s, st) in sorted(per_line.items(), key=lambda kv: -kv[1][SORT])[:top]:
    text = src[line - 1].strip()[:70] if line and line <= len(src) else ""
    top_st = ", ".join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"L{line!s:>5} inst {100*e/tot_e:5.1f}%  samp {100*s/max(tot_s,1):5.1f}%  [{top_st}]  {text}")

# ---- per-function aggregation (source line ranges) ----
import bisect
funcs = []
for i, t in enumerate(src, 1):
    m = re.match(r"(?:template\s*<[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(", t) or re.match(r"^(\w+)\(const __grid_constant__", t)
    if m and not t.strip().startswith("//"):
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
agg = collections.defaultdict(lambda: [0, 0])
for line, (e, s_, st) in per_line.items():
    if not line:
        name = "?"
    else:
        k = bisect.bisect_right(starts, line) - 1
        name = funcs[k][1] if k >= 0 else "?"
    agg[name][0] += e; agg[name][1] += s_
print("\nper function: inst%  samp%")
for name, (e, s_) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {name:24s} {100*e/tot_e:5.1f}%  {100*s_/max(tot_s,1):5.1f}%")
