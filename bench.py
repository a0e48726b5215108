#!/usr/bin/env python
"""bench.py - agent-steps/s of the fused swarm step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one env.step (one 0.1 s motion update) of every environment of the workload on random
actions that are pre-generated on the device.  Headline workload: BASELINE.json configs[2]
(SwarmACB-Foraging-v0 daisy, 16384 envs x 20 robots per GPU, weak scaling); configs[1] and [3] are
reported in ``other_workloads``.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N = 20
WORKLOADS = {
    # name: (mission, mode, envs per GPU, task id, BASELINE.json configs index)
    "foraging_daisy_16384": ("for", "daisy", 16384, "SwarmACB-Foraging-v0", 2),
    "homing_lily_4096": ("hom", "lily", 4096, "SwarmACB-Homing-v0", 1),
    "dirgate_dandelion_8192": ("dgt", "dandelion", 8192, "SwarmACB-DirectionalGate-v0", 3),
    "sheltering_oc2_16384": ("shl", "oc2", 16384, "SwarmACB-Sheltering-v0", 4),
    "xor_cyclamen_16384": ("xor", "cyclamen", 16384, "SwarmACB-XOR-v0", 0),
}
# algorithmic FLOPs and HBM bytes per agent-step
ALG_FLOPS = {"xor": 12234, "hom": 12094, "for": 12184, "dgt": 14164, "shl": 15084}


def alg_bytes_per_agent_step(discrete: bool, obs_dim: int, mission: str) -> float:
    b = 24 + 8 + 4 * obs_dim + 1.5                  # pose R+W, action, obs W, per-env scalars / 20
    if discrete:
        b += 72                                          # cached wheels 16 + fsm 8 + behaviour cache 48
    if mission == "for":
        b += 2                                           # has_food / prev_in_nest flags R+W
    return b


def make_cfg(mission, mode, E, device):
    from swarmacb_isaaclab_b200 import MISSION_CFGS
    cfg = MISSION_CFGS[mission]()
    if mode in ("oc2", "oc2c"):
        cfg.update_variant("cyclamen")
        cfg.use_continuous_actions(full_observations=(mode == "oc2"))
    else:
        cfg.update_variant(mode)
    cfg.scene.num_envs = E
    cfg.sim.device = device
    cfg.seed = 0
    return cfg


class ClockSampler:
    """nvidia-smi clock/throttle sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index: int) -> dict:
    """Multi-GPU runs: pin this rank to the host cores next to its GPU before anything is allocated, so that the
    pinned host buffers of the e2e leg are first-touched on the GPU's own NUMA node.  The node comes from sysfs
    (/sys/bus/pci/devices/<bdf>/numa_node of the GPU's PCI address); returns what was found and done."""
    info = {"numa_nodes": None, "gpu_numa_node": None, "cores_bound": 0}
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        info["numa_nodes"] = len(nodes)
        bdf = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                   # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["gpu_numa_node"] = node
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cores_bound"] = len(cpus)
    except Exception as exc:  # noqa: BLE001 - diagnostics only
        info["error"] = str(exc)[:120]
    return info


def gen_actions(torch, discrete, T, E, device, seed=1):
    g = torch.Generator(device=device).manual_seed(seed)
    if discrete:
        return torch.randint(0, 6, (T, E, N, 1), generator=g, device=device)
    return torch.rand(T, E, N, 2, generator=g, device=device) * 2 - 1


def time_steps(torch, env, actions, K, W, flush, reward_acc=None):
    """W untimed + K timed env.steps; each timed step has its own CUDA-event pair on the launch stream so
    the L2 flush between steps stays outside the timed region.  Returns per-step ms list.  ``reward_acc``: a device
    scalar that collects the team reward of every step (outside the timed regions)."""
    T = actions.shape[0]
    for w in range(W):
        _, rew, _ = env.step_tensor(actions[w % T])
        if reward_acc is not None:
            reward_acc.add_(rew.sum())
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        if flush is not None:
            flush.add_(1.0)          # 512 MB read+write > 126 MB L2
        starts[k].record()
        _, rew, _ = env.step_tensor(actions[(W + k) % T])
        stops[k].record()
        if reward_acc is not None:
            reward_acc.add_(rew.sum())
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in zip(starts, stops)]


def time_back_to_back(torch, name, device, K, instances=4):
    """The step without the ~5 us a CUDA-event pair adds around a lone launch (profiles/README.md, block times):
    ``instances`` independent batches of the workload are stepped in turn, ONE event pair around all instances * K
    launches.  Between two steps of the same batch the other batches touch more bytes than the L2 holds (the
    contract's "inputs larger than L2" alternative to flushing), so every step still finds its state in HBM.
    Secondary figure: the headline ``value`` stays the flushed, per-step-event number."""
    from swarmacb_isaaclab_b200.env import SwarmEnv
    mission, mode, E, _, _ = WORKLOADS[name]
    envs = []
    for i in range(instances):
        env = SwarmEnv(make_cfg(mission, mode, E, device), env_offset=(1000 + i) * E)
        env.reset(seed=i)
        envs.append(env)
    acts = gen_actions(torch, bool(envs[0].params.discrete_actions), 16, E, device, seed=5)
    touched = E * N * (alg_bytes_per_agent_step(bool(envs[0].params.discrete_actions), envs[0].obs_dim, mission))
    for w in range(3):
        for env in envs:
            env.step_tensor(acts[w])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for k in range(K):
        for env in envs:
            env.step_tensor(acts[k % 16])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / (K * instances)
    l2_mb = torch.cuda.get_device_properties(device).L2_cache_size / 1e6
    return {"ms_per_step": ms, "value": E * N / (ms * 1e-3), "unit": "agent-steps/s (this rank's GPU)",
            "instances": instances, "steps_per_instance": K,
            "mb_touched_between_revisits": touched * (instances - 1) / 1e6, "l2_mb": l2_mb,
            "note": "one event pair around instances * steps launches of independent batches stepped in turn; "
                    "secondary - `value` is the flushed per-step-event figure"}


def time_rollouts(torch, env, T, reps, flush):
    """Mean ms of one decision = SwarmEnv.rollout(action, T) with a fresh random action per decision (L2 flushed)."""
    E = env.num_envs
    acts = gen_actions(torch, bool(env.params.discrete_actions), 16, E, env.device, seed=3)
    for w in range(3):
        env.rollout(acts[w], T)
    ms = []
    for k in range(reps):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.rollout(acts[k % 16], T)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms)


def port_steps(mission, mode, E, K, W, min_seconds=0.0):
    """The oracle port (oracle/swarm_oracle.c, OpenMP over envs) on this host's cores: W warm-up steps (they also
    spin up the OpenMP team), then K timed steps, extended until ``min_seconds`` have passed.  Returns
    (agent-steps/s, ms per step, steps, cores)."""
    from oracle import oracle
    from swarmacb_isaaclab_b200 import build_params
    p = build_params(make_cfg(mission, mode, E, "cpu"))
    cores = oracle.set_threads()
    rng = np.random.default_rng(0)
    host = oracle.new_state(E)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=rng.random((8, E, N, 2), dtype=np.float32),
                 yaw_u=rng.random((E, N), dtype=np.float32))
    oracle.reset(p, host, **noise)
    dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
    acts = rng.integers(0, 6, (16, E, N), dtype=np.int64) if p.discrete_actions else \
        (rng.random((16, E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    for w in range(max(W, 3)):
        oracle.step(p, host, acts[w % 16], turn_dur=dur, **noise)
    t0 = time.perf_counter()
    steps = 0
    while steps < K or time.perf_counter() - t0 < min_seconds:
        oracle.step(p, host, acts[steps % 16], turn_dur=dur, **noise)
        steps += 1
    dt = time.perf_counter() - t0
    return E * N * steps / dt, dt / steps * 1e3, steps, cores


def reference_torch_probe(mission, mode, E, K, W, threads):
    """Child process of the reference-torch leg: the UNMODIFIED reference env classes (imported from /root/reference
    in the build container, from the offline install under baseline/_ref elsewhere) under the isaaclab stub of
    tests/golden/refstub.py, eager CPU torch, protocol of BASELINE.md section 3."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import torch
    import refstub
    pkg = refstub.find_reference_package()
    if pkg is None:
        print(json.dumps({"unavailable": "no reference package (neither /root/reference nor baseline/_ref)"}))
        return
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    env = refstub.make_ref_env(mission, mode, E)
    env.reset()
    gen = torch.Generator().manual_seed(1)
    agents = env.cfg.possible_agents
    T = 8
    acts = torch.randint(0, 6, (T, E, N, 1), generator=gen) if env.cfg.discrete_actions else \
        torch.rand(T, E, N, 2, generator=gen) * 2 - 1
    for w in range(W):
        env.step({a: acts[w % T][:, i] for i, a in enumerate(agents)})
    t0 = time.perf_counter()
    for k in range(K):
        env.step({a: acts[k % T][:, i] for i, a in enumerate(agents)})
    dt = time.perf_counter() - t0
    print(json.dumps({"value": E * N * K / dt, "unit": "agent-steps/s", "ms_per_step": dt / K * 1e3, "envs": E, "steps": K,
                      "warmup": W, "threads": threads, "torch": torch.__version__, "package": pkg,
                      "env_class": type(env).__name__}))


def reference_torch(mission, mode, E, K, W, threads, timeout_s=600):
    """Run reference_torch_probe in a child (its import stubs must not leak into this process)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-torch-probe", "--probe",
           json.dumps([mission, mode, E, K, W, threads])]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        line = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not line:
            return {"unavailable": f"probe failed: {(res.stderr or res.stdout)[-200:]}"}
        return json.loads(line[-1])
    except subprocess.TimeoutExpired:
        return {"unavailable": f"probe exceeded {timeout_s} s"}


def cpu_baseline(mission, mode, E):
    """cpu_baseline leg of the default run: the oracle port at the FULL workload size for >= 2 s after a team
    warm-up, and - when the reference package is importable on this box - the reference's own torch step on a
    bounded sample (1024 envs; its per-agent throughput is flat above ~1k envs), on all cores and on one."""
    value, ms, steps, cores = port_steps(mission, mode, E, 10, 3, min_seconds=4.0)
    out = {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "ms_per_step": ms,
           "sample": f"{E} envs x {steps} steps of {mission}/{mode}: oracle/swarm_oracle.c, OpenMP on {cores} threads, "
                     "noise pre-drawn, 3 warm-up steps"}
    ref_all = reference_torch(mission, mode, 1024, 24, 3, cores)
    out["reference_torch"] = ref_all if "unavailable" in ref_all else {
        **ref_all, "kind": "reference-torch",
        "sample": "1024 envs x 24 steps of the unmodified reference env class, eager CPU torch on all cores"}
    if "unavailable" not in ref_all:
        one = reference_torch(mission, mode, 1024, 6, 1, 1)
        out["reference_torch_1_thread"] = one
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  When the
    reference package is importable (baseline/_ref, installed offline by tools/install_reference.sh, travels with the
    snapshot) that is the UNMODIFIED torch env class, each step a bounded 1024-env sample of the workload (its
    per-agent throughput is flat above ~1k envs; 16384 envs would need ~25 GB of (E,N,N,S) temporaries and minutes per
    step); otherwise the C oracle port at the full workload size.  The other arm is always reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mission, mode, E_gpu, task, idx = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    cores = len(os.sched_getaffinity(0))
    ref = reference_torch(mission, mode, 1024, K, W, cores)
    pv, pms, psteps, pcores = port_steps(mission, mode, E_gpu, min(K, 50), W, min_seconds=2.0)
    port = {"value": pv, "unit": "agent-steps/s", "cores": pcores, "kind": "port", "ms_per_step": pms,
            "sample": f"{E_gpu} envs x {psteps} steps, oracle/swarm_oracle.c with OpenMP (the full workload size)"}
    if "unavailable" not in ref:
        value, ms, kind, ncores = ref["value"], ref["ms_per_step"], "reference", ref["threads"]
        sample = (f"each step = 1024 envs x 20 robots of the reference's {ref['env_class']} ({task} {mode}), eager CPU torch "
                  f"{ref['torch']}, a bounded sample of the {E_gpu}-env workload")
    else:
        value, ms, kind, ncores, sample = pv, pms, "port", pcores, port["sample"]
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{task} {mode}, {E_gpu} envs x 20 robots per GPU, random actions (BASELINE.json configs[{idx}])",
                   "reference_arm": "unmodified reference env classes (CPU torch)" if kind == "reference" else
                                    "CPU oracle port of the reference step (no reference package on this box: "
                                    + str(ref.get("unavailable")) + ")"},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": ncores, "kind": kind, "sample": sample,
                         "port": port, "reference_torch": ref},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_reset_steps(torch, env, actions, reps, flush):
    """ms of an env.step in which EVERY env times out: terminal critic snapshot, respawn (rejection sampling),
    all-env collision re-solve (ENV:1242-1273), first observation - the step BASELINE's >= 2000-step protocol crosses
    once per 1200 / 1800 steps."""
    ms = []
    for k in range(reps):
        env.episode_length_buf.fill_(env.max_episode_length - 1)
        env._check_len_buf()                    # rebuild the any-reset flag outside the timed region
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, _, to = env.step_tensor(actions[k % actions.shape[0]])
        b.record()
        torch.cuda.synchronize()
        assert bool(to.all())
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms)


def load_profile_counters():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")))["workloads"]
    except (OSError, KeyError, ValueError):
        return {}


def issue_roofline(counters, ms_step, sms, sm_mhz, peak_src):
    """The binding roofline of this kernel is the SM's instruction ISSUE rate (FP32 / ALU work on CUDA cores; no tensor
    cores, DRAM at ~5 %): one warp instruction per scheduler per cycle, 4 schedulers per SM.  achieved = executed
    warp instructions per launch (ncu counter smsp__inst_executed.sum of the committed capture of THIS workload's
    kernel) / the launch duration measured in this run; lane_weighted also discounts the inactive lanes
    (smsp__thread_inst_executed / 32).  Every figure can be recomputed from profiles/ and this line."""
    peak = sms * 4 * sm_mhz * 1e6
    if not counters:
        return {"bound": "fp32-issue", "achieved": None, "peak": peak / 1e9, "unit": "G warp-instr/s", "frac": None,
                "traffic": None, "note": "no committed ncu counters for this workload under profiles/"}
    achieved = counters["warp_instructions"] / (ms_step * 1e-3)
    frac = achieved / peak
    return {
        "bound": "fp32-issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instr/s", "frac": frac,
        "lane_weighted_frac": frac * counters["threads_per_instruction"] / 32.0,
        "peak_source": f"{sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz ({peak_src})",
        "kernel": counters["kernel"], "warp_instructions_per_launch": counters["warp_instructions"],
        "threads_per_instruction": counters["threads_per_instruction"],
        "ncu": {k: counters.get(k) for k in ("gpu_time_us", "issue_active_pct", "pipe_alu_pct", "pipe_fma_pct",
                                            "sm_cycles_active_avg", "sm_cycles_elapsed_avg", "registers_per_thread")},
        "traffic": counters.get("dram_bytes_per_launch"),
        "traffic_note": "dram__bytes_read+write of the committed ncu capture (warm L2: the working set is L2-resident "
                        "across ncu's replays), NOT measured in this run; timed steps here start from a flushed L2",
        "counters_from": counters.get("source"),
        "timing": "CUDA events around each env.step launch on the launch stream (one swarm_kernel launch per step)",
    }


def measure_workload(torch, name, device, K, W, flush, env_offset_rank, want_e2e):
    from swarmacb_isaaclab_b200 import _lib
    from swarmacb_isaaclab_b200.env import SwarmEnv
    mission, mode, E, task, idx = WORKLOADS[name]
    cfg = make_cfg(mission, mode, E, device)
    env = SwarmEnv(cfg, env_offset=env_offset_rank * E)
    env.reset(seed=0)
    discrete = bool(env.params.discrete_actions)
    T = 64
    actions = gen_actions(torch, discrete, T, E, device)
    lib = _lib.load()
    l0 = lib.swarm_kernel_launch_count()
    reward_acc = torch.zeros((), dtype=torch.float64, device=device)
    ms = time_steps(torch, env, actions, K, W, flush, reward_acc)
    launches = lib.swarm_kernel_launch_count() - l0 - W
    res = {"env": env, "E": E, "mission": mission, "mode": mode, "task": task, "idx": idx, "ms": ms,
           "launches": launches, "discrete": discrete, "reward_acc": reward_acc}
    if want_e2e:
        # end-to-end through the C ABI with HOST buffers: pinned actions H2D, step, obs/reward/time_out D2H
        h_act = actions.cpu().pin_memory()
        h_obs = torch.empty(E, N, env.obs_dim, dtype=torch.float32).pin_memory()
        h_rew = torch.empty(E, dtype=torch.float32).pin_memory()
        h_to = torch.empty(E, dtype=torch.uint8).pin_memory()

        def host_step(t):
            env.step_host(h_act[t % T], h_obs, h_rew, h_to)

        for w in range(W):
            host_step(w)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            host_step(W + k)
        torch.cuda.synchronize()
        res["e2e_s"] = time.perf_counter() - t0
        res["h2d"] = h_act[0].numel() * h_act.element_size()
        res["d2h"] = h_obs.numel() * 4 + h_rew.numel() * 4 + h_to.numel()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="foraging_daisy_16384", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-torch-probe"])
    ap.add_argument("--probe", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference-torch-probe":
        return reference_torch_probe(*json.loads(args.probe))
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the swarm step has no CPU fallback "
                         "(use --impl reference for the CPU baseline arm)")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"cores_bound": 0, "note": "N = 1 keeps every core"}
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))

    K, W = args.steps, args.warmup
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=device)  # 512 MB > L2
    from swarmacb_isaaclab_b200 import _lib
    lib = _lib.load()

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    head = measure_workload(torch, args.workload, device, K, W, flush, rank, want_e2e=True)
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()

    total_ms = sum(head["ms"])
    t = torch.tensor([total_ms, head["e2e_s"]], dtype=torch.float64, device=device)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)                    # every rank's own numbers (attribution of the max)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks
    per_rank_ms = [float(x[0]) / K for x in per_rank]
    per_rank_e2e_ms = [float(x[1]) * 1e3 / K for x in per_rank]
    total_ms, e2e_s = float(t[0]), float(t[1])
    E, mission = head["E"], head["mission"]
    agent_steps = E * N * K * world
    value = agent_steps / (total_ms * 1e-3)
    e2e_value = agent_steps / e2e_s

    # episode-metric reduction: the only collective of the path
    from swarmacb_isaaclab_b200.sharding import EpisodeMetrics
    env = head["env"]
    em = EpisodeMetrics(device)
    em.vec[3] = head["reward_acc"]          # team reward summed over every env and every step of the device-timed leg
    em.vec[4] = float(E * N * (K + W))
    metrics = em.reduce()   # NCCL all-reduce(sum) when world > 1

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    prof = load_profile_counters()
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    sm_mhz, sm_src = (float(peaks["sm_max_mhz"]), "MEASURED_PEAKS.json sm_max_mhz") if "sm_max_mhz" in peaks else \
        (float(clocks.get("sm_max_mhz") or 1965.0), "nvidia-smi clocks.max.sm")
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    ms_step = total_ms / K
    med_ms = statistics.median(head["ms"])
    bytes_as = alg_bytes_per_agent_step(head["discrete"], env.obs_dim, mission)
    alg_bytes = bytes_as * E * N
    achieved_gbs = alg_bytes / (ms_step * 1e-3) / 1e9
    fp32 = C.c_float(0.0)
    lib.swarm_fp32_peak(20000, C.byref(fp32), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    alg_flops = ALG_FLOPS[mission] * E * N
    achieved_tf = alg_flops / (ms_step * 1e-3) / 1e12
    warm = time_steps(torch, env, gen_actions(torch, head["discrete"], 8, E, device, seed=2), K, 3, None)
    reset_ms = time_reset_steps(torch, env, gen_actions(torch, head["discrete"], 4, E, device, seed=4), 5, flush)
    head_m5 = time_rollouts(torch, env, 5, min(K, 60), flush)   # the trainers' cadence: one action held for 5 steps
    back_to_back = time_back_to_back(torch, args.workload, device, min(K, 100))

    others = {}
    if not args.no_others:
        for name in ("homing_lily_4096", "dirgate_dandelion_8192", "sheltering_oc2_16384", "xor_cyclamen_16384"):
            if name == args.workload:
                continue
            r = measure_workload(torch, name, device, min(K, 100), 5, flush, 0, want_e2e=False)
            m = sum(r["ms"]) / len(r["ms"])
            others[name] = {"value": r["E"] * N / (m * 1e-3), "unit": "agent-steps/s (1 GPU)", "ms_per_step": m,
                            "config": f"BASELINE.json configs[{r['idx']}]" if r["idx"] else
                                      "XOR env of configs[0] at the headline batch size",
                            "roofline": issue_roofline(prof.get(name), m, sms, sm_mhz, sm_src)}
            # the trainers' cadence (one action held for decision_period=5 motion updates, agents/poca_trainer.py:564-573)
            # through SwarmEnv.rollout: one fused launch per decision (module actions still run the sensors every step)
            m5 = time_rollouts(torch, r["env"], 5, min(K, 60), flush)
            others[name]["decision_period_5"] = {"value": r["E"] * N * 5 / (m5 * 1e-3), "unit": "agent-steps/s (1 GPU)",
                                                 "ms_per_decision": m5,
                                                 "path": "swarm_rollout, fused kernel"}
            if prof.get(name + "@rollout5"):
                others[name]["decision_period_5"]["roofline"] = issue_roofline(prof[name + "@rollout5"], m5, sms, sm_mhz, sm_src)
            if name == "sheltering_oc2_16384":
                # BASELINE.json configs[4] without its trainer (the reference's PyTorch code, not shipped here): the
                # env side of the OC2 loop at the trainer's cadence - per decision a fresh on-device action sample,
                # get_critic_state() and one rollout of 5 motion updates - wall clock through the Python API
                from swarmacb_isaaclab_b200 import runner
                rr = runner.random_policy_rollout(r["env"], decisions=min(K, 100), decision_period=5, seed=5)
                others[name]["trainer_cadence_wall_clock"] = {
                    "value": rr["agent_steps_per_s"], "unit": "agent-steps/s (1 GPU)",
                    "agent_decisions_per_s": rr["agent_decisions_per_s"], "decisions": rr["decisions"],
                    "path": "runner.random_policy_rollout: torch action sample + get_critic_state + SwarmEnv.rollout(5)"}

        # BASELINE.json configs[0]: the manual_control.py kinematic path, 1 env x 20 robots, one 180 s episode
        # (1800 ticks of MC:721-757), wall clock through StandaloneSwarmEnv.tick (launch-latency bound at E = 1)
        from swarmacb_isaaclab_b200.standalone import StandaloneSwarmEnv
        mc = StandaloneSwarmEnv(20, device, "SwarmACB-XOR-v0", num_envs=1, seed=0)
        ids = torch.randint(0, 6, (64, 1, N), device=device)
        for w in range(20):
            mc.tick(ids[w % 64])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(1800):
            mc.tick(ids[k % 64])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        others["manual_control_xor_1env"] = {"value": N * 1800 / dt, "unit": "agent-steps/s (1 GPU, wall clock)",
                                             "ms_per_step": dt / 1800 * 1e3,
                                             "config": "BASELINE.json configs[0]: 1 env x 20 e-pucks, 1800-tick rollout"}

    cpu = None if args.no_cpu else cpu_baseline(mission, head["mode"], E)

    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{head['task']} {head['mode']}, {E} envs x 20 robots per GPU, random module actions "
                        f"(BASELINE.json configs[{head['idx']}])",
            "envs_per_gpu": E, "robots_per_env": N, "decimation": 1, "noise": "in-kernel Philox4x32-10 (2 blocks per robot-step)",
            "l2": "512 MB buffer rewritten between timed steps (L2 flushed); per-step CUDA events",
            "parallelism": f"env-sharded x{world}, no collective in the step",
            "host_cores_bound_per_rank": numa.get("cores_bound", 0), "host_numa": numa,
        },
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": head["h2d"],
                "d2h_bytes_per_step": head["d2h"],
                "path": "SwarmEnv.step_host -> swarm_host_step (C ABI): pinned host actions H2D, step, obs+reward+time_out D2H "
                        "into pinned host buffers, env chunks pipelined over a copy stream, stream sync",
                "bound": "PCIe: the observation download alone is d2h_bytes_per_step / (PCIe D2H rate) per step"},
        "gpu_launches": head["launches"],
        "roofline": issue_roofline(prof.get(args.workload), ms_step, sms, sm_mhz, sm_src),
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_agent_step": bytes_as,
                         "note": "not the binding bound: the step is issue-bound, its DRAM throughput is ~5 % of peak"},
        "roofline_fp32_algorithmic": {
            "bound": "fp32-issue (algorithmic FLOP count)", "achieved": achieved_tf, "peak": float(fp32.value),
            "unit": "TFLOP/s", "frac": achieved_tf / float(fp32.value) if fp32.value > 0 else None,
            "algorithmic_flops_per_agent_step": ALG_FLOPS[mission],
            "peak_source": "swarm_fp32_peak FMA micro-benchmark, this run",
            "note": "SECONDARY: SURVEY 8d's count prices every pair / ray / segment test the reference's maths implies; "
                    "the kernel's exact culling skips most of them, so this is not a utilisation figure"},
        "per_rank": {"ms_per_step": per_rank_ms, "e2e_ms_per_step": per_rank_e2e_ms},
        "reset_step_ms": reset_ms,
        "back_to_back": back_to_back,
        "decision_period_5": {"value": E * N * 5 * world / (head_m5 * 1e-3) if world == 1 else None,
                              "unit": "agent-steps/s (this rank's GPU)", "ms_per_decision": head_m5,
                              "path": "SwarmEnv.rollout -> swarm_rollout, one fused launch per 5-step decision "
                                      "(agents/poca_trainer.py:564-573)",
                              "roofline": issue_roofline(prof.get(args.workload + "@rollout5"), head_m5, sms, sm_mhz, sm_src)},
        "ms_per_step_median": med_ms, "ms_per_step_warm_l2": sum(warm) / len(warm),
        "clocks": clocks,
        "episode_metrics": {"sum_group_reward": metrics["sum_group_reward"], "agent_steps": metrics["agent_steps"],
                            "reduced_over_ranks": world},
        "other_workloads": others,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
