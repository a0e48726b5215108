#!/usr/bin/env python
"""List the backward branches (loops) of one kernel with their span in instructions / bytes and source line.

    python tools/sass_loops.py lib.so 'swarm_kernelILi3ELb1ELi24ELi0E' [ncu_source_page.csv]

A hot loop whose body exceeds the ~6 KB L0 instruction cache refetches every iteration (profiling aid).
"""
import csv, os, re, subprocess, sys, tempfile

so, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
execd = {}
if len(sys.argv) > 3:
    rows = list(csv.reader(open(sys.argv[3])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ia, ie = h.index("Address"), h.index("Instructions Executed")
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    base = int(body[0][ia], 16)
    for r in body:
        execd[int(r[ia], 16) - base] = int(r[ie] or 0)
active, labels, instrs, chain, cur = False, {}, [], [], None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        active = pat in m.group(1)
        continue
    if not active:
        continue
    m = re.match(r"^(\.L_x_\d+):", ln)
    if m:
        labels[m.group(1)] = len(instrs)
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if chain:
            cur = next((c for c in chain if c[0] == "swarm_step.cu" and c[1] > 90), chain[0])
        chain = []
        instrs.append((int(m.group(1), 16), m.group(2), cur))
print(f"{len(instrs)} instructions")
out = []
for k, (addr, text, line) in enumerate(instrs):
    m = re.search(r"\bBRA\S*\s+.*?`\((\.L_x_\d+)\)", text)
    if m and m.group(1) in labels and labels[m.group(1)] <= k:
        t = labels[m.group(1)]
        out.append((k - t + 1, t, k, line, execd.get(addr, 0)))
for span, t, k, line, ex in sorted(out, reverse=True):
    print(f"span {span:5d} instr ({span*16/1024:5.1f} KB)  [{t:5d}..{k:5d}]  back-branch at line {line}  executed {ex}")
