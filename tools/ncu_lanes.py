#!/usr/bin/env python
"""Lane occupancy per CUDA source region from an ncu SASS source page: executed warp instructions, average active
threads and the issue slots lost to inactive lanes ((32 - avg threads) / 32 x instructions).

    python tools/ncu_lanes.py sass.csv lib.so 'swarm_kernelILi3ELb1ELi24ELi0E' [top]
"""
import bisect, collections, csv, os, re, subprocess, sys, tempfile

sass_csv, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, chain, active, cur = {}, [], False, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        active = pat in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m:
        if chain:
            cur = next((c for c in chain if c[0] == "swarm_step.cu" and c[1] > 115), chain[0])
        chain = []
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) == len(h)]
ia, ie, it = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
base = int(body[0][ia], 16)
src_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "swarmacb-isaaclab_b200", "csrc", "swarm_step.cu")
src = open(src_path).read().splitlines()
per = collections.defaultdict(lambda: [0, 0])
for r in body:
    key = line_of.get(int(r[ia], 16) - base)
    e, t = int(r[ie] or 0), int(r[it] or 0)
    per[key][0] += e
    per[key][1] += t
tot_e = sum(v[0] for v in per.values())
tot_t = sum(v[1] for v in per.values())
print(f"warp instr {tot_e}, avg threads {tot_t / tot_e:.2f}, lost issue slots {100 * (1 - tot_t / (32 * tot_e)):.1f}%")
funcs = []
for i, t in enumerate(src, 1):
    m = re.match(r"(?:template\s*<[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(", t) or re.match(r"^(\w+)\(const __grid_constant__", t)
    if m and not t.strip().startswith("//"):
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
agg = collections.defaultdict(lambda: [0, 0])
for key, (e, t) in per.items():
    if key is None or key[0] != "swarm_step.cu":
        name = "other:" + (key[0] if key else "?")
    else:
        k = bisect.bisect_right(starts, key[1]) - 1
        name = funcs[k][1] if k >= 0 else "?"
    agg[name][0] += e
    agg[name][1] += t
print("function                  inst%   avg-threads   lost-slots% (of all issue slots)")
for name, (e, t) in sorted(agg.items(), key=lambda kv: -(kv[1][0] - kv[1][1] / 32)):
    if e:
        print(f"  {name:24s} {100 * e / tot_e:5.1f}   {t / e:6.1f}       {100 * (e - t / 32) / tot_e:5.1f}")
print("lines with the most lost slots:")
for key, (e, t) in sorted(per.items(), key=lambda kv: -(kv[1][0] - kv[1][1] / 32))[:top]:
    text = src[key[1] - 1].strip()[:74] if key and key[0] == "swarm_step.cu" and key[1] <= len(src) else ""
    print(f"  {str(key[1] if key else None):>5} inst {100 * e / tot_e:4.1f}%  thr {t / max(e, 1):5.1f}  lost {100 * (e - t / 32) / tot_e:4.1f}%  {text}")
