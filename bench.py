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


def bind_to_gpu_numa_node(index: int) -> int:
    """Multi-GPU runs: pin this rank to the host cores next to its GPU before anything is allocated, so that the
    pinned host buffers of the e2e leg are first-touched on the GPU's own NUMA node (8 ranks x 31.5 MB of
    observations per step otherwise cross the socket interconnect).  Returns the number of cores bound (0 = left
    alone: NVML unavailable or nothing to restrict)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, (n + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1 and i in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def gen_actions(torch, discrete, T, E, device, seed=1):
    g = torch.Generator(device=device).manual_seed(seed)
    if discrete:
        return torch.randint(0, 6, (T, E, N, 1), generator=g, device=device)
    return torch.rand(T, E, N, 2, generator=g, device=device) * 2 - 1


def time_steps(torch, env, actions, K, W, flush):
    """W untimed + K timed env.steps; each timed step has its own CUDA-event pair on the launch stream so
    the L2 flush between steps stays outside the timed region.  Returns per-step ms list."""
    T = actions.shape[0]
    for w in range(W):
        env.step_tensor(actions[w % T])
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        if flush is not None:
            flush.add_(1.0)          # 512 MB read+write > 126 MB L2
        starts[k].record()
        env.step_tensor(actions[(W + k) % T])
        stops[k].record()
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in zip(starts, stops)]


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


def cpu_baseline(mission, mode, budget_s=12.0, E=1024, threads=None):
    """The oracle port (oracle/swarm_oracle.c, OpenMP over envs) timed on this host's cores on a bounded
    sample of the same workload: E envs x as many steps as fit the budget."""
    from oracle import oracle
    from swarmacb_isaaclab_b200 import build_params
    cfg = make_cfg(mission, mode, E, "cpu")
    p = build_params(cfg)
    cores = oracle.set_threads()
    rng = np.random.default_rng(0)
    host = oracle.new_state(E)
    oracle.reset(p, host, rab_u=rng.random((E, N, N), dtype=np.float32),
                 spawn_u=rng.random((8, E, N, 2), dtype=np.float32), yaw_u=rng.random((E, N), dtype=np.float32))
    rab_u = rng.random((E, N, N), dtype=np.float32)
    dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
    spawn_u, yaw_u = rng.random((8, E, N, 2), dtype=np.float32), rng.random((E, N), dtype=np.float32)
    if p.discrete_actions:
        acts = rng.integers(0, 6, (16, E, N), dtype=np.int64)
    else:
        acts = (rng.random((16, E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    for w in range(3):
        oracle.step(p, host, acts[w], rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
    t0 = time.perf_counter()
    steps = 0
    while True:
        oracle.step(p, host, acts[steps % 16], rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
        steps += 1
        if time.perf_counter() - t0 > budget_s or steps >= 4000:
            break
    dt = time.perf_counter() - t0
    return {"value": E * N * steps / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"{E} envs x {steps} steps of {mission}/{mode} (oracle/swarm_oracle.c, OpenMP, noise pre-drawn)",
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric = the oracle port on all host threads
    (the reference itself is Python under /root/reference, which does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mission, mode, E_gpu, task, idx = WORKLOADS[args.workload]
    E = 1024
    from oracle import oracle
    from swarmacb_isaaclab_b200 import build_params
    cfg = make_cfg(mission, mode, E, "cpu")
    p = build_params(cfg)
    cores = oracle.set_threads()
    rng = np.random.default_rng(0)
    host = oracle.new_state(E)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=rng.random((8, E, N, 2), dtype=np.float32),
                 yaw_u=rng.random((E, N), dtype=np.float32))
    oracle.reset(p, host, **noise)
    dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
    acts = rng.integers(0, 6, (16, E, N), dtype=np.int64) if p.discrete_actions else \
        (rng.random((16, E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    K, W = args.steps, args.warmup
    for w in range(W):
        oracle.step(p, host, acts[w % 16], turn_dur=dur, **noise)
    t0 = time.perf_counter()
    for k in range(K):
        oracle.step(p, host, acts[(W + k) % 16], turn_dur=dur, **noise)
    dt = time.perf_counter() - t0
    value = E * N * K / dt
    sample = f"each step = {E} envs x 20 robots of {task} {mode} (bounded sample of the {E_gpu}-env workload)"
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{task} {mode}, {E_gpu} envs x 20 robots per GPU, random actions (BASELINE.json configs[{idx}])",
                   "reference_arm": "CPU oracle port of the reference step (reference is pure Python/torch and cannot travel)"},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    ms = time_steps(torch, env, actions, K, W, flush)
    launches = lib.swarm_kernel_launch_count() - l0 - W
    res = {"env": env, "E": E, "mission": mission, "mode": mode, "task": task, "idx": idx, "ms": ms,
           "launches": launches, "discrete": discrete}
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
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
    bound_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else 0   # N = 1 keeps every core (cpu_baseline leg)
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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks
    total_ms, e2e_s = float(t[0]), float(t[1])
    E, mission = head["E"], head["mission"]
    agent_steps = E * N * K * world
    value = agent_steps / (total_ms * 1e-3)
    e2e_value = agent_steps / e2e_s

    # episode-metric reduction: the only collective of the path
    from swarmacb_isaaclab_b200.sharding import EpisodeMetrics
    env = head["env"]
    em = EpisodeMetrics(device)
    em.vec[3] = env._episode_group_reward.sum().double() + env.completed_group_reward.sum().double()
    em.vec[4] = float(E * N * (K + W))
    metrics = em.reduce()   # NCCL all-reduce(sum) when world > 1

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, prof = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    try:  # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_metrics.json")))
    except OSError:
        pass
    traffic = prof.get("dram_bytes_per_launch") if args.workload == "foraging_daisy_16384" else None
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

    others = {}
    if not args.no_others:
        for name in ("homing_lily_4096", "dirgate_dandelion_8192", "sheltering_oc2_16384", "xor_cyclamen_16384"):
            if name == args.workload:
                continue
            r = measure_workload(torch, name, device, min(K, 100), 5, flush, 0, want_e2e=False)
            m = sum(r["ms"]) / len(r["ms"])
            others[name] = {"value": r["E"] * N / (m * 1e-3), "unit": "agent-steps/s (1 GPU)", "ms_per_step": m,
                            "config": f"BASELINE.json configs[{r['idx']}]" if r["idx"] else
                                      "XOR env of configs[0] at the headline batch size"}
            # the trainers' cadence (one action held for decision_period=5 motion updates, agents/poca_trainer.py:564-573)
            # through SwarmEnv.rollout: one fused launch for wheel actions, 5 back-to-back launches for module actions
            m5 = time_rollouts(torch, r["env"], 5, min(K, 60), flush)
            others[name]["decision_period_5"] = {"value": r["E"] * N * 5 / (m5 * 1e-3), "unit": "agent-steps/s (1 GPU)",
                                                 "ms_per_decision": m5,
                                                 "path": "swarm_rollout, fused kernel" if not r["discrete"] else
                                                         "swarm_rollout, 5 launches"}
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

    cpu = None if args.no_cpu else cpu_baseline(mission, head["mode"])

    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{head['task']} {head['mode']}, {E} envs x 20 robots per GPU, random module actions "
                        f"(BASELINE.json configs[{head['idx']}])",
            "envs_per_gpu": E, "robots_per_env": N, "decimation": 1, "noise": "in-kernel Philox4x32-10",
            "l2": "512 MB buffer rewritten between timed steps (L2 flushed); per-step CUDA events",
            "parallelism": f"env-sharded x{world}, no collective in the step",
            "host_cores_bound_per_rank": bound_cores,
        },
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": head["h2d"],
                "d2h_bytes_per_step": head["d2h"],
                "path": "SwarmEnv.step_host -> swarm_host_step (C ABI): pinned host actions H2D, step, obs+reward+time_out D2H "
                        "into pinned host buffers, 8 env chunks pipelined over a copy stream, stream sync"},
        "gpu_launches": head["launches"],
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": "swarm_kernel<FOR,discrete,24,STEP>" if args.workload == "foraging_daisy_16384" else "swarm_kernel",
                     "algorithmic_bytes_per_agent_step": bytes_as,
                     "timing": "CUDA events around each env.step launch (one swarm_kernel launch per step)"},
        "roofline_fp32": {"bound": "fp32-issue", "achieved": achieved_tf, "peak": float(fp32.value), "unit": "TFLOP/s",
                          "frac": achieved_tf / float(fp32.value) if fp32.value > 0 else None,
                          "algorithmic_flops_per_agent_step": ALG_FLOPS[mission],
                          "peak_source": "swarm_fp32_peak FMA micro-benchmark, this run"},
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
