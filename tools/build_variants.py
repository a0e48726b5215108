"""Tuning aid: build several -D variants of the kernel library in parallel (tuning log)."""
import sys, os, concurrent.futures as cf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swarmacb_isaaclab_b200 import build as B
VARIANTS = dict(a.split("=", 1) for a in sys.argv[1:])  # name=DEF1,DEF2
out_dir = os.path.join(B.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
def one(item):
    name, defs = item
    return B.build_variant(os.path.join(out_dir, f"lib_{name}.so"), [d for d in defs.split(",") if d])
with cf.ThreadPoolExecutor(8) as ex:
    for r in ex.map(one, VARIANTS.items()):
        print("built", r)
