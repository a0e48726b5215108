"""GPU parity at scale: CUDA step vs the CPU oracle on seeded random rollouts (teacher-forced per step),
plus size-independent properties at BASELINE.json's full sizes."""
import zlib

import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle
from swarmacb_isaaclab_b200.params import N

pytestmark = pytest.mark.gpu

CASES = [
    ("hom", "lily", 1), ("for", "daisy", 1), ("dgt", "dandelion", 1), ("shl", "oc2", 1),
    ("xor", "cyclamen", 1), ("shl", "daisy", 2), ("dgt", "daisy", 1), ("xor", "oc2c", 1),
]


def _close(name, got, want, tol, lab):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    d = np.abs(got - want)
    if not d.max() <= tol:
        idx = np.unravel_index(d.argmax(), d.shape)
        bad = np.argwhere(d > tol)
        raise AssertionError(f"{lab}: {name} max|diff|={d.max():.3e} > {tol:g} at {idx} got={got[idx]!r} "
                             f"want={want[idx]!r}; {len(bad)} elements over tol, cols={sorted(set(bad[:, -1].tolist()))[:24]}")


def _mk(mission, mode, E, dec=1):
    from swarmacb_isaaclab_b200.env import SwarmEnv
    return SwarmEnv(fixtures.make_cfg(mission, mode, E, dec, device="cuda:0"))


def _cluster(rng, E, frac=0.5):
    """Half of the envs get a tight robot cluster so that collisions / rays / RAB are busy."""
    spawn = rng.random((6, E, N, 2), dtype=np.float32)
    k = int(E * frac)
    c = rng.random((k, 1, 2), dtype=np.float32) * 0.8 + 0.1
    spawn[:, :k] = c + (spawn[:, :k] - 0.5) * 0.12
    return spawn


@pytest.mark.parametrize("mission,mode,dec", CASES)
def test_cuda_vs_oracle_rollout(mission, mode, dec):
    E, T = 384, 12
    rng = np.random.default_rng(zlib.crc32(f"{mission}/{mode}".encode()))  # stable across processes (str hash is salted)
    env = _mk(mission, mode, E, dec)
    p = env.params
    host = oracle.new_state(E)
    spawn_u, yaw_u = _cluster(rng, E), rng.random((E, N), dtype=np.float32)
    rab_u = rng.random((E, N, N), dtype=np.float32)
    env.inject_noise(rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    env.reset()
    obs_o = oracle.reset(p, host, rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    torch.cuda.synchronize()
    dev = env.dump_state()
    assert np.array_equal(dev["pos"], host["pos"]) and np.array_equal(dev["yaw"], host["yaw"])
    _close("reset obs", env._obs.cpu().numpy(), obs_o, 0.0, f"{mission}/{mode} reset")
    # push some envs close to the time limit so that partial resets happen inside the window
    host["episode_length_buf"][::7] = p.max_episode_length - 5
    host["episode_length_buf"][3::11] = p.max_episode_length - 9
    for t in range(T):
        env.load_state(host)  # teacher-forced: both sides start every step from the oracle's state
        if p.discrete_actions:
            act = rng.integers(0, 6, (E, N), dtype=np.int64)
        else:
            act = (rng.random((E, N, 2), dtype=np.float32) * 2.4 - 1.2).astype(np.float32)
        rab_u = rng.random((E, N, N), dtype=np.float32)
        dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
        spawn_u, yaw_u = _cluster(rng, E), rng.random((E, N), dtype=np.float32)
        env.inject_noise(rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
        obs, rew, to = env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        obs_o, rew_o, to_o = oracle.step(p, host, act, rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
        crit = env.get_critic_state().cpu().numpy()
        torch.cuda.synchronize()
        dev = env.dump_state()
        lab = f"{mission}/{mode} t={t}"
        # the pose path is built from exactly rounded ops on both sides: poses must agree bit for bit
        assert np.array_equal(dev["pos"], host["pos"]), f"{lab}: pos not bit-identical, max diff {np.abs(dev['pos'] - host['pos']).max():.3e}"
        assert np.array_equal(dev["yaw"], host["yaw"]), f"{lab}: yaw not bit-identical"
        assert np.array_equal(dev["cached_left"], host["cached_left"]), f"{lab}: wheels not bit-identical"
        for k in ("fsm", "mission_flags", "episode_length_buf", "prev_ground", "episode_group_reward",
                  "completed_group_reward"):
            assert np.array_equal(dev[k], host[k]), f"{lab}: {k}"
        assert np.array_equal(rew.cpu().numpy(), rew_o), lab
        assert np.array_equal(to.cpu().numpy(), to_o), lab
        _close("obs", obs.cpu().numpy(), obs_o, 0.0, lab)        # same float32 op sequence on both sides: bit-exact
        if p.discrete_actions:
            _close("beh_cache", dev["beh_cache"], host["beh_cache"], 0.0, lab)
        assert np.abs(crit - oracle.critic_state(p, host)).max() <= 2e-5, lab
        assert np.abs(dev["completed_terminal_critic_state"] - host["completed_terminal_critic_state"]).max() <= 2e-5


@pytest.mark.parametrize("mission,mode", [("shl", "daisy"), ("for", "oc2"), ("dgt", "cyclamen")])
def test_crowded_queues_overflow_into_further_rounds(mission, mode):
    """The sensor suite's per-warp work queues hold 64 / 64 / 32 items; whatever does not fit goes through another
    round.  Here every packet survives (loss probability 0) and all 20 robots of every env sit in one 12-cm cluster:
    ~600 range-and-bearing items and several hundred ray-disc items per warp, i.e. about ten rounds of every queue.
    Results must still equal the oracle bit for bit."""
    E = 96
    rng = np.random.default_rng(21)
    cfg = fixtures.make_cfg(mission, mode, E, device="cuda:0")
    cfg.rab_loss_probability = 0.0
    from swarmacb_isaaclab_b200.env import SwarmEnv
    env = SwarmEnv(cfg)
    p = env.params
    host = oracle.new_state(E)
    spawn_u, yaw_u = _cluster(rng, E, frac=1.0), rng.random((E, N), dtype=np.float32)
    rab_u = rng.random((E, N, N), dtype=np.float32)
    env.inject_noise(rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    env.reset()
    obs_o = oracle.reset(p, host, rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    torch.cuda.synchronize()
    assert np.array_equal(env._obs.cpu().numpy(), obs_o)
    ztilde = obs_o[..., 19 if p.obs_dim == 24 else 3]
    assert ztilde.mean() > 0.99          # every robot hears (almost) all 19 others
    for t in range(3):
        env.load_state(host)
        act = rng.integers(0, 6, (E, N), dtype=np.int64) if p.discrete_actions else \
            (rng.random((E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
        rab_u = rng.random((E, N, N), dtype=np.float32)
        dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
        env.inject_noise(rab_u=rab_u, turn_dur=dur)
        obs, rew, _ = env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        obs_o, rew_o, _ = oracle.step(p, host, act, rab_u=rab_u, turn_dur=dur)
        dev = env.dump_state()
        assert np.array_equal(dev["pos"], host["pos"]) and np.array_equal(dev["yaw"], host["yaw"]), t
        assert np.array_equal(obs.cpu().numpy(), obs_o), t
        assert np.array_equal(rew.cpu().numpy(), rew_o), t


@pytest.mark.parametrize("mission,mode,E", [("hom", "lily", 4096), ("for", "daisy", 16384), ("dgt", "dandelion", 8192)])
def test_full_size_properties(mission, mode, E):
    """BASELINE-size rollouts with in-kernel Philox noise: invariants that need no oracle."""
    env = _mk(mission, mode, E)
    p = env.params
    env.reset(seed=3)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    total = torch.zeros(E, device="cuda:0")
    for t in range(30):
        if p.discrete_actions:
            act = torch.randint(0, 6, (E, N, 1), generator=g, device="cuda:0")
        else:
            act = torch.rand(E, N, 2, generator=g, device="cuda:0") * 2 - 1
        obs, rew, to = env.step_tensor(act)
        total += rew
    torch.cuda.synchronize()
    pos = env.agent_pos
    assert torch.isfinite(pos).all() and torch.isfinite(obs).all()
    inr = 1.2357309
    assert (pos.norm(dim=-1) <= inr / np.cos(np.pi / 12) + 1e-3).all()          # inside the circumcircle
    d = (pos.unsqueeze(2) - pos.unsqueeze(1)).norm(dim=-1) + torch.eye(N, device=pos.device) * 10
    assert (d.min() > 0.03)                                                       # solver keeps robots apart
    assert (env.agent_yaw.abs() <= np.pi + 1e-6).all()
    assert (env.episode_length_buf == 30).all()
    assert torch.equal(rew, rew.round()) and (rew >= -N).all() and (rew <= N).all()  # integer-valued counters
    assert torch.allclose(env._episode_group_reward, total)
    zt = obs[..., 19] if p.obs_dim == 24 else obs[..., 3]
    assert (zt >= 0).all() and (zt < 1).all()
    # packet loss p=0.85: mean neighbour count must be far below the no-loss count
    keep_rate = float((zt > 0).float().mean())
    assert 0.02 < keep_rate < 0.9
    # determinism: same seed, same actions -> identical trajectory; different shard offset -> different noise
    env2 = _mk(mission, mode, E)
    env2.reset(seed=3)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for t in range(30):
        if p.discrete_actions:
            act = torch.randint(0, 6, (E, N, 1), generator=g, device="cuda:0")
        else:
            act = torch.rand(E, N, 2, generator=g, device="cuda:0") * 2 - 1
        obs2, _, _ = env2.step_tensor(act)
    assert torch.equal(env2.agent_pos, env.agent_pos) and torch.equal(obs2, obs)


def _sharded_setup(episode_length_s=None):
    from swarmacb_isaaclab_b200.env import SwarmEnv
    from swarmacb_isaaclab_b200.sharding import shard_cfg
    base = fixtures.make_cfg("for", "daisy", 96, device="cuda:0")
    base.seed = 11
    if episode_length_s is not None:
        base.episode_length_s = episode_length_s
    whole = SwarmEnv(base)
    parts = []
    for r in range(3):
        c, off = shard_cfg(base, 96, 3, r, device="cuda:0")
        parts.append(SwarmEnv(c, env_offset=off))
    whole.reset()
    for p_ in parts:
        p_.reset()
    return whole, parts


def _sharded_run(whole, parts, steps, expect_equal=True):
    g = torch.Generator(device="cuda:0").manual_seed(5)
    rollovers, equal = 0, True
    for t in range(steps):
        act = torch.randint(0, 6, (96, N, 1), generator=g, device="cuda:0")
        obs, rew, to = whole.step_tensor(act)
        outs = [p_.step_tensor(act[32 * r: 32 * (r + 1)].contiguous()) for r, p_ in enumerate(parts)]
        rollovers += int(to.sum())
        # counters, rewards before any divergence and time-out flags never depend on the re-solve coupling
        assert torch.equal(to, torch.cat([o[2] for o in outs])), t
        same = torch.equal(obs, torch.cat([o[0] for o in outs])) and torch.equal(rew, torch.cat([o[1] for o in outs]))
        if expect_equal:
            assert same, f"step {t}"
        equal = equal and same
    same_pos = torch.equal(whole.agent_pos, torch.cat([p_.agent_pos for p_ in parts]))
    if expect_equal:
        assert same_pos
    return rollovers, equal and same_pos


def test_sharded_rollout_equals_unsharded():
    """the Philox stream is keyed by the global env index, so cutting a job into shards
    (one per GPU) reproduces the unsharded trajectory bit for bit - across synchronised roll-overs too
    (1 s episodes = 10 steps: every env of every shard respawns and is re-solved at the same steps)."""
    whole, parts = _sharded_setup()
    _sharded_run(whole, parts, 25)
    whole, parts = _sharded_setup(episode_length_s=1.0)
    rollovers, _ = _sharded_run(whole, parts, 34)
    assert rollovers == 3 * 96


def test_sharded_rollout_with_desynchronised_counters():
    """ENV:1262 re-solves ALL envs of the batch whenever ANY env times out.  With episode counters out of lockstep the
    batch-wide flag differs between a shard and the whole job; the JobResetClock hands every shard the job-wide
    flag (one all-reduce when it is built, none per step) and restores bit-identity.  Crowded spawn (Homing strip)
    so that the extra re-solve actually moves robots."""
    from swarmacb_isaaclab_b200.env import SwarmEnv
    from swarmacb_isaaclab_b200.sharding import shard_cfg

    def setup(with_clock):
        base = fixtures.make_cfg("hom", "lily", 96, device="cuda:0")
        base.seed, base.episode_length_s = 4, 2.0          # 20-step episodes
        whole = SwarmEnv(base)
        parts = []
        for r in range(3):
            c, off = shard_cfg(base, 96, 3, r, device="cuda:0")
            parts.append(SwarmEnv(c, env_offset=off))
        for env in [whole] + parts:
            env.reset()
        lens = torch.zeros(96, dtype=torch.long)
        lens[:32] = torch.arange(32) % 20                    # only shard 0 is staggered: it rolls over at every step,
        whole.episode_length_buf = lens                      # shards 1 and 2 only every 20th
        for r, p_ in enumerate(parts):
            p_.episode_length_buf = lens[32 * r: 32 * (r + 1)]
        if with_clock:
            for p_ in parts:
                p_.attach_job_reset_clock(peers=parts)
        return whole, parts

    rollovers, equal = _sharded_run(*setup(False), 45, expect_equal=False)
    assert rollovers > 96 and not equal       # per-shard coupling (the default) is visible once counters are staggered
    _sharded_run(*setup(True), 45)            # job-wide flags: bit-identical again


def test_dict_api_and_zero_copy_actions():
    """The reference protocol: dict of 20 strided views in, dicts of views out (poca_trainer.py:559-573)."""
    env = _mk("xor", "cyclamen", 64)
    obs, info = env.reset()
    assert set(obs) == {f"epuck_{i}" for i in range(N)} and obs["epuck_3"].shape == (64, 4)
    actions = torch.randint(0, 6, (64, N, 1), device="cuda:0")
    act_dict = {a: actions[:, i] for i, a in enumerate(env.cfg.possible_agents)}
    assert env._gather_actions(act_dict).data_ptr() == actions.data_ptr()     # zero-copy path
    obs, rew, term, trunc, info = env.step(act_dict)
    a0 = env.cfg.possible_agents[0]
    assert rew[a0].shape == (64,) and trunc[a0].dtype == torch.bool and not term[a0].any()
    assert obs["epuck_19"].shape == (64, 4)
    # separate (non-view) tensors and int32 ids go through the copy path and give the same result
    env2 = _mk("xor", "cyclamen", 64)
    env2.reset()
    act_dict2 = {a: actions[:, i].clone().to(torch.int32) for i, a in enumerate(env.cfg.possible_agents)}
    obs2, rew2, *_ = env2.step(act_dict2)
    assert torch.equal(obs2[a0], obs[a0]) and torch.equal(rew2[a0], rew[a0])
    assert env.unwrapped is env and env.get_critic_state().shape == (64, N, 5)
    assert env.max_episode_length == 1800 and env.episode_length_buf.dtype == torch.long


@pytest.mark.parametrize("mission,mode", [("hom", "lily"), ("shl", "daisy"), ("dgt", "dandelion")])
def test_free_running_timeouts_match_oracle(mission, mode):
    """No teacher forcing: 14 consecutive steps with 5-step episodes, so the kernel's own rotating any-reset
    flags (and an external write to episode_length_buf that de-synchronises the envs) drive full and partial
    resets, including the all-env re-solve of ENV:1262.  Poses must stay bit-identical to the oracle."""
    from swarmacb_isaaclab_b200.env import SwarmEnv
    E = 48
    cfg = fixtures.make_cfg(mission, mode, E, device="cuda:0")
    cfg.episode_length_s = 0.5                       # 5 steps per episode
    env = SwarmEnv(cfg)
    p = env.params
    assert p.max_episode_length == 5
    rng = np.random.default_rng(7)
    host = oracle.new_state(E)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=_cluster(rng, E), yaw_u=rng.random((E, N), dtype=np.float32))
    env.inject_noise(**noise)
    env.reset()
    oracle.reset(p, host, **noise)
    resets = 0
    for t in range(14):
        if t == 2:  # de-synchronise: from now on envs 0..15 time out two steps earlier than the rest
            env.episode_length_buf[:16] += 2
            host["episode_length_buf"][:16] += 2
        act = rng.integers(0, 6, (E, N), dtype=np.int64) if p.discrete_actions else \
            (rng.random((E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
        noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), turn_dur=rng.integers(1, 5, (E, N, 3)).astype(np.int32),
                     spawn_u=_cluster(rng, E), yaw_u=rng.random((E, N), dtype=np.float32))
        env.inject_noise(**noise)
        obs, rew, to = env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        obs_o, rew_o, to_o = oracle.step(p, host, act, **noise)
        torch.cuda.synchronize()
        dev = env.dump_state()
        resets += int(to_o.sum())
        assert np.array_equal(to.cpu().numpy(), to_o), t
        assert np.array_equal(dev["pos"], host["pos"]), f"t={t}: max diff {np.abs(dev['pos'] - host['pos']).max():.3e}"
        assert np.array_equal(dev["episode_length_buf"], host["episode_length_buf"])
        assert np.array_equal(rew.cpu().numpy(), rew_o)
        _close("obs", obs.cpu().numpy(), obs_o, 0.0, f"{mission}/{mode} t={t}")
        if p.discrete_actions:
            assert np.array_equal(dev["fsm"], host["fsm"]) and np.array_equal(dev["beh_cache"], host["beh_cache"])
    assert resets >= 2 * E


@pytest.mark.parametrize("E", [1, 5, 8, 9, 17, 33])
def test_ragged_batch_sizes(E):
    """Batches that do not fill the last 8-environment block (its spare slots shadow the last env without stores)."""
    env = _mk("shl", "daisy", E)
    p = env.params
    rng = np.random.default_rng(E)
    host = oracle.new_state(E)
    guard = torch.full((E + 2, N, 24), 7.0, device="cuda:0")          # canaries around the obs buffer
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=_cluster(rng, E), yaw_u=rng.random((E, N), dtype=np.float32))
    env.inject_noise(**noise)
    env.reset()
    oracle.reset(p, host, **noise)
    for t in range(4):
        act = rng.integers(0, 6, (E, N), dtype=np.int64)
        noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), turn_dur=rng.integers(1, 5, (E, N, 3)).astype(np.int32))
        env.inject_noise(**noise)
        obs, rew, to = env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        obs_o, rew_o, _ = oracle.step(p, host, act, **noise)
        torch.cuda.synchronize()
        dev = env.dump_state()
        assert np.array_equal(dev["pos"], host["pos"]) and np.array_equal(obs.cpu().numpy(), obs_o)
        assert np.array_equal(rew.cpu().numpy(), rew_o)
    assert (guard == 7.0).all()


@pytest.mark.parametrize("mission,mode,E", [("for", "daisy", 8192 + 48), ("dgt", "dandelion", 5000), ("hom", "lily", 1000)])
def test_host_buffer_step_equals_device_step(mission, mode, E):
    """swarm_host_step (pinned host buffers, chunked H2D / step / D2H pipeline) must give exactly what swarm_step
    gives on the same seed: chunking only re-partitions independent environments."""
    a, b = _mk(mission, mode, E), _mk(mission, mode, E)
    a.reset(seed=3)
    b.reset(seed=3)
    # some environments roll over inside the window (partial resets + the all-env re-solve)
    for env in (a, b):
        env.episode_length_buf[::5] = env.max_episode_length - 3
    g = torch.Generator().manual_seed(5)
    discrete = bool(a.params.discrete_actions)
    h_obs = torch.empty(E, N, a.obs_dim).pin_memory()
    h_rew = torch.empty(E).pin_memory()
    h_to = torch.empty(E, dtype=torch.uint8).pin_memory()
    rolled = 0
    for t in range(6):
        if discrete:
            act = torch.randint(0, 6, (E, N, 1), generator=g)
        else:
            act = torch.rand(E, N, 2, generator=g) * 2 - 1
        h_act = act.pin_memory()
        obs, rew, to = a.step_tensor(act.to("cuda:0"))
        b.step_host(h_act, h_obs, h_rew, h_to)
        assert torch.equal(obs.cpu(), h_obs), f"t={t}: observations differ"
        assert torch.equal(rew.cpu(), h_rew) and torch.equal(to.cpu().to(torch.uint8), h_to), f"t={t}"
        rolled += int(h_to.sum())
    sa, sb = a.dump_state(), b.dump_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert rolled == len(range(0, E, 5))


@pytest.mark.parametrize("mission,mode,dec,T,repeat", [
    ("for", "daisy", 1, 5, True), ("for", "daisy", 1, 7, False), ("hom", "lily", 1, 5, True),
    ("dgt", "dandelion", 1, 5, True), ("shl", "oc2", 1, 37, False), ("xor", "cyclamen", 2, 6, False),
    ("shl", "oc2c", 1, 5, True), ("dgt", "daisy", 1, 33, True),
])
@pytest.mark.parametrize("fused", [True, False])
def test_fused_rollout_equals_single_steps(mission, mode, dec, T, repeat, fused, monkeypatch):
    """swarm_rollout (T env.steps fused in one launch, state in registers, sensors skipped where nothing reads
    them) must leave exactly the state, last observation, summed reward and OR-ed time_out of T swarm_step
    calls on the same Philox stream - including envs that roll over inside the window, the all-env re-solve
    they trigger, and the single steps that follow the rollout (rotating any-reset flags rebuilt)."""
    if fused:           # the default: one fused launch per <= 32 steps, for wheel and module actions alike
        monkeypatch.delenv("SWARM_UNFUSED_ROLLOUT", raising=False)
    else:               # the fallback path: back-to-back single-step launches
        monkeypatch.setenv("SWARM_UNFUSED_ROLLOUT", "1")
    E = 600
    a, b = _mk(mission, mode, E, dec), _mk(mission, mode, E, dec)
    a.reset(seed=11)
    b.reset(seed=11)
    L = a.max_episode_length
    for env in (a, b):   # roll-overs at different steps of the window; some envs untouched
        env.episode_length_buf[::7] = L - 2
        env.episode_length_buf[3::13] = L - 4
        env.episode_length_buf[5::17] = L - 1
    g = torch.Generator(device="cuda:0").manual_seed(9)
    discrete = bool(a.params.discrete_actions)
    n_act = 1 if repeat else T
    if discrete:
        acts = torch.randint(0, 6, (n_act, E, N, 1), generator=g, device="cuda:0")
    else:
        acts = torch.rand(n_act, E, N, 2, generator=g, device="cuda:0") * 2.4 - 1.2
    rew = torch.zeros(E, device="cuda:0")
    to = torch.zeros(E, dtype=torch.bool, device="cuda:0")
    for t in range(T):
        obs_a, r, d = a.step_tensor(acts[0 if repeat else t])
        rew += r
        to |= d
    obs_b, rew_b, to_b = b.rollout(acts[0] if repeat else acts, T)
    assert torch.equal(obs_a, obs_b), "last observation differs"
    assert torch.equal(rew, rew_b) and torch.equal(to, to_b)
    assert int(to.sum()) >= len(range(0, E, 7))
    sa, sb = a.dump_state(), b.dump_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert torch.equal(a.get_critic_state(), b.get_critic_state())
    # single steps after the rollout: flags, counters and Philox stream continue identically
    b.episode_length_buf[1::19] = L - 2
    a.episode_length_buf[1::19] = L - 2
    for t in range(3):
        oa, ra, da = a.step_tensor(acts[0])
        ob, rb, db = b.step_tensor(acts[0])
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), f"post-rollout step {t}"
    sa, sb = a.dump_state(), b.dump_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k


def test_headless_runner_random_policy(tmp_path, capsys):
    """`python -m swarmacb_isaaclab_b200.runner --config X.yaml` without a reference checkout: random policy at the
    trainers' cadence through SwarmEnv.rollout, one JSON line of throughput and episode metrics."""
    import json
    from swarmacb_isaaclab_b200 import runner
    path = tmp_path / "run.yaml"
    path.write_text("behaviors:\n  DirGate_dandelion:\n    task: SwarmACB-DirectionalGate-v0\n    variant: dandelion\n"
                    "    environment:\n      num_envs: 256\n      decision_period: 5\n      episode_length_s: 3.0\n")
    assert runner.main(["--config", str(path), "--decisions", "12", "--seed", "2"]) == 0
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert out["envs"] == 256 and out["env_steps"] == 60 and out["obs_dim"] == 24 and not out["discrete_actions"]
    assert out["episodes_finished"] == 256 * 2 and out["agent_steps_per_s"] > 0


@pytest.mark.parametrize("mission,mode", [("for", "daisy"), ("shl", "oc2")])
def test_checkpoint_resume_is_bit_exact(mission, mode, tmp_path):
    """state_dict() -> torch.save -> a fresh env's load_state_dict(): the resumed run continues bit for bit
    (poses, FSM state, counters, Philox position), across a roll-over."""
    E = 300
    a = _mk(mission, mode, E)
    a.reset(seed=4)
    a.episode_length_buf[::3] = a.max_episode_length - 6
    g = torch.Generator(device="cuda:0").manual_seed(1)
    discrete = bool(a.params.discrete_actions)
    acts = (torch.randint(0, 6, (12, E, N, 1), generator=g, device="cuda:0") if discrete
            else torch.rand(12, E, N, 2, generator=g, device="cuda:0") * 2 - 1)
    for t in range(4):
        a.step_tensor(acts[t])
    torch.save(a.state_dict(), tmp_path / "env.pt")
    b = _mk(mission, mode, E)
    b.load_state_dict(torch.load(tmp_path / "env.pt"))
    assert torch.equal(a._obs, b._obs)
    rolled = 0
    for t in range(4, 12):
        oa, ra, da = a.step_tensor(acts[t])
        ob, rb, db = b.step_tensor(acts[t])
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), f"t={t}"
        rolled += int(da.sum())
    assert rolled == len(range(0, E, 3))
    sa, sb = a.dump_state(), b.dump_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    with pytest.raises(ValueError):
        _mk(mission, mode, E + 1).load_state_dict(a.state_dict())


def test_philox_packet_loss_statistics():
    """Production noise: every ordered in-range pair keeps its packet with P = 0.15 (SENS:419-421), independently.
    Homing has no internal walls and robots never touch the arena faces, so line of sight never blocks and the
    neighbour count n_i = popcount(kept & in_range) can be read back from ztilde = 1 - 2 / (1 + e^n)."""
    E = 4096
    env = _mk("hom", "lily", E)
    env.reset(seed=7)
    g = torch.Generator(device="cuda:0").manual_seed(2)
    lut = torch.tensor(list(env.params.ztilde_lut), device="cuda:0")
    kept = in_range = 0
    var_num = var_den = 0.0
    for t in range(12):
        obs, _, _ = env.step_tensor(torch.randint(0, 6, (E, N, 1), generator=g, device="cuda:0"))
        n = (obs[..., 3].unsqueeze(-1) - lut).abs().argmin(-1)                       # (E,N) kept neighbour count
        pos = env.agent_pos
        d = (pos.unsqueeze(2) - pos.unsqueeze(1)).pow(2).sum(-1).add(1e-8).sqrt()
        m = ((d < 0.6) & ~torch.eye(N, dtype=torch.bool, device="cuda:0")).sum(-1)  # (E,N) neighbours in range
        kept += int(n.sum())
        in_range += int(m.sum())
        # binomial dispersion: sum (n - m p)^2 ~ sum m p (1 - p)
        var_num += float(((n - 0.15 * m).double() ** 2).sum())
        var_den += float((m * 0.15 * 0.85).double().sum())
    rate = kept / in_range
    sigma = (0.15 * 0.85 / in_range) ** 0.5
    assert in_range > 2_000_000 and abs(rate - 0.15) < 5 * sigma + 1e-4, (rate, sigma)
    assert abs(var_num / var_den - 1.0) < 0.02, var_num / var_den                  # no over/under-dispersion


@pytest.mark.parametrize("mission,mode,E", [("hom", "lily", 4096), ("for", "daisy", 16384), ("dgt", "dandelion", 8192),
                                            ("shl", "oc2", 16384)])
def test_baseline_size_batches_equal_oracle(mission, mode, E):
    """BASELINE.json configs[1..4] at their full per-GPU batch sizes: reset + 3 free-running steps (one with a
    partial roll-over) against the oracle on the same injected noise - everything bit for bit."""
    rng = np.random.default_rng(E + len(mission))
    env = _mk(mission, mode, E)
    p = env.params
    oracle.set_threads()
    host = oracle.new_state(E)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=rng.random((6, E, N, 2), dtype=np.float32),
                 yaw_u=rng.random((E, N), dtype=np.float32))
    env.inject_noise(**noise)
    env.reset()
    obs_o = oracle.reset(p, host, **noise)
    assert np.array_equal(env._obs.cpu().numpy(), obs_o)
    env.episode_length_buf[::9] = p.max_episode_length - 2
    host["episode_length_buf"][::9] = p.max_episode_length - 2
    for t in range(3):
        act = rng.integers(0, 6, (E, N), dtype=np.int64) if p.discrete_actions else \
            (rng.random((E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
        noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), turn_dur=rng.integers(1, 5, (E, N, 3)).astype(np.int32),
                     spawn_u=rng.random((6, E, N, 2), dtype=np.float32), yaw_u=rng.random((E, N), dtype=np.float32))
        env.inject_noise(**noise)
        obs, rew, to = env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        obs_o, rew_o, to_o = oracle.step(p, host, act, **noise)
        dev = env.dump_state()
        for k in ("pos", "yaw", "fsm", "prev_ground", "mission_flags", "episode_length_buf", "episode_group_reward",
                  "completed_group_reward", "cached_left", "cached_right"):
            assert np.array_equal(dev[k], host[k]), f"{mission}/{mode} E={E} t={t}: {k}"
        assert np.array_equal(obs.cpu().numpy(), obs_o) and np.array_equal(rew.cpu().numpy(), rew_o)
        assert np.array_equal(to.cpu().numpy(), to_o)
    assert host["completed_terminal_critic_state"].any()
