#!/usr/bin/env python
"""Static code-size breakdown of one kernel: SASS instructions per CUDA source function / line.

    python tools/sass_size.py swarmacb-isaaclab_b200/libswarmstep.so 'swarm_kernelILi3ELb1ELi24ELi0E' [top_lines]

(profiling aid: the kernel is instruction-fetch sensitive, so static size per phase matters)
"""
import bisect, collections, os, re, subprocess, sys, tempfile

so, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
HELPER_MAX = int(os.environ.get("HELPER_MAX", "0"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
src_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "swarmacb-isaaclab_b200", "csrc", "swarm_step.cu")
src = open(src_path).read().splitlines()
per_line = collections.Counter()
ops = collections.Counter()
active, chain, cur = False, [], None
n = 0
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        active = pat in m.group(1)
        cur = None
        continue
    if not active:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?(\S+)", ln)
    if m:
        if chain:
            cur = chain[0]
            if HELPER_MAX:   # attribute tiny helpers (fadd/fdiv/... defined above line HELPER_MAX) to their call site
                for c in chain:
                    if c[0] == "swarm_step.cu" and c[1] > HELPER_MAX:
                        cur = c
                        break
        chain = []
        per_line[cur] += 1
        ops[m.group(2).split(".")[0]] += 1
        n += 1
funcs = []
for i, t in enumerate(src, 1):
    m = re.match(r"(?:template\s*<[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(", t) or re.match(r"^(\w+)\(const __grid_constant__", t)
    if m and not t.strip().startswith("//"):
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
agg = collections.Counter()
for key, c in per_line.items():
    if key is None:
        agg["?"] += c
    elif key[0] != "swarm_step.cu":
        agg[key[0]] += c
    else:
        k = bisect.bisect_right(starts, key[1]) - 1
        agg[funcs[k][1] if k >= 0 else "?"] += c
print(f"{n} SASS instructions = {n * 16 / 1024:.1f} KB")
for name, c in agg.most_common():
    print(f"  {name:28s} {c:5d}  {100 * c / n:5.1f}%")
print("top lines:")
for key, c in per_line.most_common(top):
    text = src[key[1] - 1].strip()[:80] if key and key[0] == "swarm_step.cu" and key[1] <= len(src) else ""
    print(f"  {key}  {c:4d}  {text}")
print("top opcodes:", ", ".join(f"{k}:{v}" for k, v in ops.most_common(16)))
