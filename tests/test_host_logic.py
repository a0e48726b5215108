"""CPU-only checks: cfg mirror, params derivation, fsm packing, C-ABI exports, no-fallback behaviour."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

import swarmacb_isaaclab_b200 as pkg
from swarmacb_isaaclab_b200 import build as cuda_build
from swarmacb_isaaclab_b200 import params as P
from swarmacb_isaaclab_b200.cfg import TASK_CFGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_task_ids_and_cfg_entry_points():
    assert set(TASK_CFGS) == {
        "SwarmACB-DirectionalGate-v0", "SwarmACB-XOR-v0", "SwarmACB-Homing-v0", "SwarmACB-Foraging-v0",
        "SwarmACB-Sheltering-v0", "SwarmACB-SCA-v0", "SwarmACB-SHL-v0"}
    from swarmacb_isaaclab_b200 import env
    for tid, spec in env.registry.items():
        mod, cls = spec["kwargs"]["env_cfg_entry_point"].split(":")
        assert getattr(__import__(mod, fromlist=[cls]), cls) is TASK_CFGS[tid]


def test_variants_and_spaces():
    cfg = pkg.DirectionalGateEnvCfg()
    assert cfg.num_agents == 20 and cfg.possible_agents[0] == "epuck_0" and not cfg.discrete_actions
    for v, (od, ad) in {"dandelion": (24, 2), "daisy": (24, 1), "lily": (4, 1), "tulip": (4, 1), "cyclamen": (4, 1)}.items():
        cfg.update_variant(v)
        assert cfg.observation_spaces["epuck_19"] == od and cfg.action_spaces["epuck_0"] == ad
        assert cfg.discrete_actions == (v != "dandelion")
        assert P.build_params(cfg).obs_dim == od
    cfg.update_variant("cyclamen")
    cfg.use_continuous_actions(full_observations=True)
    p = P.build_params(cfg)
    assert (p.obs_dim, p.discrete_actions) == (24, 0) and cfg.action_spaces["epuck_3"] == 2
    cfg.use_continuous_actions(full_observations=False)
    assert P.build_params(cfg).obs_dim == 4


def test_episode_lengths_and_constants():
    want = {"dgt": 1200, "xor": 1800, "hom": 1200, "for": 1800, "shl": 1800}
    for m, cls in pkg.MISSION_CFGS.items():
        cfg = cls()
        p = P.build_params(cfg)
        assert p.max_episode_length == want[m]
        assert p.n_segments == 12 + p.n_internal
        assert abs(p.wall_r_eff - 0.0401) < 1e-7 and abs(p.two_radius - 0.07) < 1e-8
        assert math.isclose(cfg.arena_circumradius, math.sqrt(4.91 / 3), rel_tol=1e-12)
    assert P.build_params(pkg.ShelteringEnvCfg()).capsule_clearance == pytest.approx(0.0501)
    assert P.build_params(pkg.DirectionalGateEnvCfg()).capsule_clearance == pytest.approx(0.0401)
    assert P.build_params(pkg.XorAggregationEnvCfg()).gate_mode == P.GATE_DGT  # inherited push-out quirk
    cfg = pkg.DirectionalGateEnvCfg()
    cfg.decimation = 6
    assert P.build_params(cfg).max_episode_length == 200


def test_sensor_tables_are_float32_torch_results():
    cos_a, sin_a, rc, rs = P.sensor_tables()
    assert cos_a.dtype == np.float32 and cos_a[2] != 0.0 and abs(cos_a[2]) < 1e-7  # cos(float32(pi/2))
    assert sin_a[2] == -1.0 and rc[0] == pytest.approx(math.sqrt(0.5), abs=1e-7)


def test_fsm_pack_roundtrip():
    g = torch.Generator().manual_seed(0)
    E, N = 5, 20
    f = dict(
        es=torch.randint(0, 2, (E, N), generator=g), est=torch.randint(0, 5, (E, N), generator=g),
        ed=torch.randint(-1, 2, (E, N), generator=g).float(), pa=torch.randint(0, 2, (E, N), generator=g).bool(),
        ps=torch.randint(0, 5, (E, N), generator=g), pd=torch.randint(-1, 2, (E, N), generator=g).float(),
        aa=torch.randint(0, 2, (E, N), generator=g).bool(), as_=torch.randint(0, 5, (E, N), generator=g),
        ad=torch.randint(-1, 2, (E, N), generator=g).float())
    w = P.pack_fsm(*f.values())
    u = P.unpack_fsm(w)
    for got, want in zip(u.values(), f.values()):
        assert torch.equal(got.float(), want.float())


def test_abi_header_matches_ctypes_layout():
    hdr = open(os.path.join(ROOT, "include", "swarm_abi.h")).read()
    for struct, cls in (("SwarmParams", P.SwarmParams), ("SwarmState", P.SwarmState), ("SwarmNoise", P.SwarmNoise),
                        ("SwarmOut", P.SwarmOut)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*?\]", "", part.strip().split()[-1].lstrip("*")))
        assert names == [f[0] for f in cls._fields_], struct


def test_shared_library_exports_every_declared_symbol():
    from swarmacb_isaaclab_b200 import _lib
    path = cuda_build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "swarm_abi.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*) (swarm_\w+)\(", hdr, re.M))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.swarm_abi_version.restype = ctypes.c_int
    assert lib.swarm_abi_version() == P.ABI_VERSION


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from swarmacb_isaaclab_b200.env import SwarmEnv, make
    cfg = pkg.HomingEnvCfg()
    cfg.sim.device = "cpu"
    with pytest.raises(RuntimeError):
        SwarmEnv(cfg)
    with pytest.raises(RuntimeError):
        make("SwarmACB-Homing-v0")
    with pytest.raises(KeyError):
        make("SwarmACB-Nope-v0")


def test_product_never_imports_oracle():
    """The product path must not import, link or dlopen anything under oracle/ (comments may name it)."""
    src_dir = os.path.join(ROOT, "swarmacb-isaaclab_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|swarm_oracle|liboracle|oracle/_|#include.*oracle", re.M)
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def _lib_no_gpu():
    from swarmacb_isaaclab_b200 import _lib
    return _lib.load()


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Argument validation runs before any CUDA call, so the error contract is testable without a GPU:
    0 ok, <0 SWARM_E_* (-> ValueError in the Python mirror), >0 cudaError_t (-> RuntimeError)."""
    lib = _lib_no_gpu()
    p = P.build_params(pkg.HomingEnvCfg())
    st, nz, out = P.SwarmState(), P.SwarmNoise(), P.SwarmOut()
    C = ctypes
    assert lib.swarm_step(None, C.byref(st), None, C.byref(nz), C.byref(out), 4, None) == -1          # SWARM_E_NULL
    assert lib.swarm_step(C.byref(p), C.byref(st), None, C.byref(nz), C.byref(out), 0, None) == -3    # SWARM_E_SIZE
    bad = P.build_params(pkg.HomingEnvCfg())
    bad.abi_version = 99
    assert lib.swarm_step(C.byref(bad), C.byref(st), None, C.byref(nz), C.byref(out), 4, None) == -4  # SWARM_E_VERSION
    bad = P.build_params(pkg.HomingEnvCfg())
    bad.obs_dim = 7
    assert lib.swarm_step(C.byref(bad), C.byref(st), None, C.byref(nz), C.byref(out), 4, None) == -2  # SWARM_E_PARAM
    bad = P.build_params(pkg.HomingEnvCfg())
    bad.gate_mode = P.GATE_SHL
    assert lib.swarm_reset(C.byref(bad), C.byref(st), C.byref(nz), C.byref(out), 4, None) == -2
    assert lib.swarm_step(C.byref(p), C.byref(st), None, C.byref(nz), C.byref(out), 4, None) == -1    # null state ptrs
    assert b"null" in lib.swarm_last_error_string()
    assert lib.swarm_critic_state(C.byref(p), C.byref(st), None, 4, None) == -1
    assert lib.swarm_mc_tick(C.byref(p), C.byref(st), None, None, C.byref(nz), C.byref(out), 7, 1, None) == -2  # not MC params
    mc = P.build_mc_params("SwarmACB-XOR-v0")
    assert lib.swarm_mc_tick(C.byref(mc), C.byref(st), None, None, C.byref(nz), C.byref(out), 0, 1, None) == -2  # empty flags
    assert lib.swarm_rollout(C.byref(p), C.byref(st), None, 0, C.byref(nz), C.byref(out), 4, 5, None) == -1
    from swarmacb_isaaclab_b200 import _lib
    with pytest.raises(ValueError):
        _lib.check(-2, "x")
    with pytest.raises(RuntimeError):
        _lib.check(700, "x")


def test_mc_params_reproduce_reference_quirks():
    mc = P.build_mc_params("SwarmACB-XOR-v0")
    assert mc.mc_mode == 1 and mc.gate_mode == P.GATE_NONE and mc.max_episode_length == 1800
    assert mc.light_y == pytest.approx(-1.4)
    # manual_control.py:536-544: the 12th face's mid-angle wraps to pi -> duplicate of the west face
    assert mc.mc_face_nx[11] == pytest.approx(1.0) and mc.mc_face_px[11] == pytest.approx(mc.mc_face_px[5])
    assert P.build_mc_params("SwarmACB-Foraging-v0").zone[6] == pytest.approx(-0.63)
    assert P.build_mc_params("SwarmACB-Homing-v0").mc_spawn_theta_max == pytest.approx(math.pi)
    assert P.build_mc_params("nonsense").mission == P.MISSION_ID["dgt"]


_RUN_YAML = """
behaviors:
  OC2_Sheltering_cyclamen:
    task: SwarmACB-Sheltering-v0
    variant: cyclamen
    trainer_type: oc2
    max_steps: 1000
    time_horizon: 40
    environment:
      num_envs: 7
      decision_period: 4
      episode_length_s: 12.0
      no_such_key: 1
"""


def test_runner_builds_env_cfg_like_train_py(tmp_path):
    """runner.load_run_spec / build_env_cfg follow agents/config_loader.py:30-187 and scripts/train.py:166-185."""
    from swarmacb_isaaclab_b200 import ShelteringEnvCfg, runner
    path = tmp_path / "run.yaml"
    path.write_text(_RUN_YAML)
    spec = runner.load_run_spec(str(path))
    assert (spec.run_name, spec.task_id, spec.variant) == ("OC2_Sheltering_cyclamen", "SwarmACB-Sheltering-v0", "cyclamen")
    assert spec.trainer_type == "learned_option_critic" and spec.decision_period == 4
    assert spec.env_overrides == {"num_envs": 7, "episode_length_s": 12.0, "no_such_key": 1}
    warned = []
    cfg = runner.build_env_cfg(spec.task_id, spec.variant, spec.trainer_type, spec.env_overrides, seed=3,
                               device="cuda:0", warn=warned.append)
    assert isinstance(cfg, ShelteringEnvCfg) and cfg.scene.num_envs == 7 and cfg.episode_length_s == 12.0
    assert cfg.seed == 3 and not cfg.discrete_actions and cfg.full_policy_observations   # CFG:195-209
    assert len(warned) == 1 and "no_such_key" in warned[0]
    with pytest.raises(KeyError):
        runner.build_env_cfg("SwarmACB-Nope-v0", "lily")
    with pytest.raises(FileNotFoundError):
        runner.load_run_spec(str(tmp_path / "missing.yaml"))
    with pytest.raises(FileNotFoundError):
        runner.import_reference_agents(str(tmp_path))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0
    # "reference" = the unmodified torch env classes (when the reference package is importable: /root/reference here,
    # baseline/_ref on the GPU box), else the C oracle port; the other arm is reported beside it
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 0 and "16384 envs" in cb["port"]["sample"]
    if cb["kind"] == "reference":
        assert cb["reference_torch"]["env_class"] == "ForagingEnv" and cb["reference_torch"]["envs"] == 1024
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_rollout_rejects_ambiguous_or_short_action_buffers():
    """swarm_rollout cannot bounds-check the action buffer, so SwarmEnv.rollout validates shapes itself
    (host logic, exercised on CPU tensors with the oracle behind the C-ABI call sites)."""
    import pytest
    import fixtures
    from oracle_env import OracleBackedEnv
    E = 4
    for mission, mode, A, dt in (("xor", "cyclamen", 1, torch.long), ("dgt", "dandelion", 2, torch.float32)):
        env = OracleBackedEnv(fixtures.make_cfg(mission, mode, E))
        env.reset()
        one = torch.zeros(E, P.N, A, dtype=dt)
        per_step = torch.zeros(3, E, P.N, A, dtype=dt)
        env.rollout(one, 2)                      # one action, repeated
        env.rollout(per_step)                    # one action per step
        env.rollout(per_step, 3)
        with pytest.raises(ValueError):
            env.rollout(one)                     # repeated action without `steps`
        with pytest.raises(ValueError):
            env.rollout(per_step, 5)             # more steps than actions: would read past the buffer
        with pytest.raises(ValueError):
            env.rollout(torch.zeros(E, E, P.N, A, dtype=dt)[:, :2], 2)   # wrong env count
        with pytest.raises(ValueError):
            env.rollout(one, 0)
    # (T, E, N) integer ids are accepted for module actions; a first dimension that happens to equal E is still per-step
    env = OracleBackedEnv(fixtures.make_cfg("xor", "cyclamen", E))
    env.reset()
    env.rollout(torch.zeros(E, E, P.N, dtype=torch.long))
    assert int(env.episode_length_buf[0]) == E
    # a job-wide reset clock hands the kernel a 32-bit schedule: longer calls are refused BEFORE any state advances
    env.attach_job_reset_clock()
    counter = env._step_counter
    with pytest.raises(ValueError):
        env.rollout(torch.zeros(E, P.N, 1, dtype=torch.long), 33)
    assert env._step_counter == counter
    env.rollout(torch.zeros(E, P.N, 1, dtype=torch.long), 32)


def test_state_attribute_assignment_copies_into_the_kernel_tensor():
    import fixtures
    from oracle_env import OracleBackedEnv
    env = OracleBackedEnv(fixtures.make_cfg("hom", "lily", 3))
    env.reset()
    ptr = env.episode_length_buf.data_ptr()
    env.episode_length_buf = torch.tensor([5, 6, 7])
    assert env.episode_length_buf.data_ptr() == ptr and env.episode_length_buf.tolist() == [5, 6, 7]
    env.agent_pos = torch.zeros(3, P.N, 2)
    assert float(env.agent_pos.abs().max()) == 0.0


def test_gymnasium_registration_round_trip():
    """missions/*/__init__.py + scripts/train.py:96-106,188 of the reference: the seven ids are registered with the
    ``env_cfg_entry_point`` kwarg, the cfg class is resolved from ``gym.spec(id).kwargs`` and the env is built with
    ``gym.make(id, cfg=cfg)``.  gymnasium is not installed here, so its registration API is played by tests/gym_stub.py;
    on this CPU-only box the resolved entry point must be SwarmEnv refusing to run without CUDA."""
    import importlib
    import gym_stub
    from swarmacb_isaaclab_b200 import env as env_mod
    gym = gym_stub.install()
    try:
        assert env_mod.register_gym() is True
        assert set(gym.registry) >= set(TASK_CFGS)
        for tid, cfg_cls in TASK_CFGS.items():
            spec = gym.spec(tid)
            mod, _, name = spec.kwargs["env_cfg_entry_point"].partition(":")
            assert getattr(importlib.import_module(mod), name) is cfg_cls          # scripts/train.py:96-106
            assert spec.extra.get("disable_env_checker") is True
        cfg = TASK_CFGS["SwarmACB-XOR-v0"]()
        cfg.update_variant("cyclamen")
        cfg.scene.num_envs = 2
        if not torch.cuda.is_available():
            cfg.sim.device = "cpu"
            with pytest.raises(RuntimeError, match="CUDA"):
                gym.make("SwarmACB-XOR-v0", cfg=cfg)                                 # scripts/train.py:188
        # with gymnasium importable the per-agent spaces are gymnasium objects
        from oracle_env import OracleBackedEnv
        e = OracleBackedEnv(cfg)
        assert isinstance(e.observation_space("epuck_0"), gym.spaces.Box) and e.observation_space("epuck_0").shape == (4,)
        assert isinstance(e.action_space("epuck_3"), gym.spaces.Discrete) and e.action_space("epuck_3").n == 6
    finally:
        gym_stub.uninstall()


def test_viewer_feed_scene_frame_and_svg():
    """the viewer feed - static scene from the SwarmParams block, one env's dynamic state, an SVG picture.
    (host logic; device state is served by the oracle on CPU here, the GPU twin is in tests/test_gpu_protocol.py)"""
    import fixtures
    from oracle_env import OracleBackedEnv
    from swarmacb_isaaclab_b200 import viewer
    for mission, mode, n_walls, n_zones in (("shl", "daisy", 3, 3), ("dgt", "dandelion", 2, 2), ("for", "cyclamen", 0, 3)):
        env = OracleBackedEnv(fixtures.make_cfg(mission, mode, 3))
        env.reset()
        sc = viewer.scene(env)
        assert sc["arena_faces"].shape == (12, 4) and sc["internal_walls"].shape == (n_walls, 4) and len(sc["zones"]) == n_zones
        fr = viewer.frame(env, 2)
        assert fr["pos"].shape == (P.N, 2) and fr["obs"].shape == (P.N, env.obs_dim) and fr["episode_step"] == 0
        assert np.array_equal(fr["pos"], env.agent_pos[2].numpy())
        assert set(fr["behaviour"]) >= {"_explore_state", "_photo_avoiding", "_antiphoto_steps"}
        svg = viewer.render_svg(sc, fr, selected=4)
        assert svg.startswith("<svg") and svg.count('class="robot"') == P.N and svg.rstrip().endswith("</svg>")
    with pytest.raises(IndexError):
        viewer.frame(env, 3)


def test_arena_face_tables_are_centrally_symmetric():
    """The kernel's conservative face screening (cand_faces, the sensor suite's band tasks) gets all twelve signed
    distances from six dot products: it relies on the regular dodecagon's central symmetry in the SwarmParams tables
    (ENV:849-872: inward normal = -midpoint/|midpoint|, point = midpoint)."""
    for m, cls in pkg.MISSION_CFGS.items():
        p = P.build_params(cls())
        inr = math.hypot(p.face_px[0], p.face_py[0])
        for f in range(6):
            assert abs(p.face_nx[f] + p.face_nx[f + 6]) < 1e-6 and abs(p.face_ny[f] + p.face_ny[f + 6]) < 1e-6, (m, f)
        for f in range(12):
            assert abs(p.face_px[f] * p.face_nx[f] + p.face_py[f] * p.face_ny[f] + inr) < 1e-6, (m, f)
            assert abs(math.hypot(p.face_px[f], p.face_py[f]) - inr) < 1e-6
