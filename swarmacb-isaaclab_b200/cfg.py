"""Environment configuration objects for the five SwarmACB missions.

Host-side mirror of the reference's ``@configclass`` env cfgs (same attribute names, defaults and
methods) so ``scripts/train.py:166-185``-style code - ``update_variant``,
``use_continuous_actions``, ``cfg.scene.num_envs = n``, ``setattr`` of YAML ``environment:`` keys -
works unchanged:

* base / DirectionalGate: directional_gate_env_cfg.py:76-209
* XOR: xor_aggregation_env_cfg.py:14-25, Homing: homing_env_cfg.py:14-25,
  Foraging: foraging_env_cfg.py:14-28, Sheltering: sheltering_env_cfg.py:14-31

No Isaac Lab dependency: ``SimCfg`` / ``SceneCfg`` only carry the fields the step path reads
(``sim.dt``, ``sim.device``, ``scene.num_envs``).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field

NUM_AGENTS = 20
ARENA_N_SIDES = 12
ARENA_AREA = 4.91
ARENA_CIRCUMRADIUS = math.sqrt(2 * ARENA_AREA / (ARENA_N_SIDES * math.sin(2 * math.pi / ARENA_N_SIDES)))

OBS_DIM = {"dandelion": 24, "daisy": 24, "lily": 4, "tulip": 4, "cyclamen": 4}
ACT_DIM = {"dandelion": 2, "daisy": 1, "lily": 1, "tulip": 1, "cyclamen": 1}
NUM_BEHAVIOR_MODULES = 6


def agent_names(n: int = NUM_AGENTS) -> list[str]:
    return [f"epuck_{i}" for i in range(n)]


def _spaces(dim: int, n: int) -> dict[str, int]:
    return {f"epuck_{i}": dim for i in range(n)}


@dataclass
class SimCfg:
    dt: float = 0.1
    render_interval: int = 1
    gravity: tuple = (0.0, 0.0, -9.81)
    device: str = "cuda:0"


@dataclass
class SceneCfg:
    num_envs: int = 5
    env_spacing: float = 4.0
    replicate_physics: bool = True


@dataclass
class DirectionalGateEnvCfg:
    """Directional-gate mission; also the base class of the other four (CFG:76-209)."""

    seed: int | None = None
    variant: str = "dandelion"
    num_agents: int = NUM_AGENTS
    possible_agents: list = field(default_factory=agent_names)
    observation_spaces: dict = field(default_factory=lambda: _spaces(24, NUM_AGENTS))
    action_spaces: dict = field(default_factory=lambda: _spaces(2, NUM_AGENTS))
    state_space: int = -1
    discrete_actions: bool = False
    num_actions: int = NUM_BEHAVIOR_MODULES
    full_policy_observations: bool = False

    decimation: int = 1
    episode_length_s: float = 120.0
    sim: SimCfg = field(default_factory=SimCfg)
    scene: SceneCfg = field(default_factory=SceneCfg)

    arena_num_sides: int = ARENA_N_SIDES
    arena_area: float = ARENA_AREA
    arena_circumradius: float = ARENA_CIRCUMRADIUS
    critic_state_radius: float = 1.20
    arena_wall_height: float = 0.08
    arena_wall_thickness: float = 0.01

    robot_radius: float = 0.035
    robot_height: float = 0.05
    robot_mass: float = 0.190
    max_wheel_speed: float = 0.16
    wheelbase: float = 0.055
    collision_solver_iterations: int = 4
    wall_contact_epsilon: float = 1e-4
    internal_wall_thickness: float = 0.01

    prox_range: float = 0.10
    rab_range: float = 0.60
    rab_loss_probability: float = 0.85
    unity_unit_scale_m: float = 0.10
    light_threshold: float = 0.2
    light_intensity: float = 1000.0

    spawn_area_center: tuple = (0.0, 0.0)
    spawn_area_size: tuple = (2.4, 2.4)
    spawn_circle_radius: float = 1.2
    spawn_max_attempts: int = 100

    debug_visual_sensors: bool = False
    sensor_visual_robot_index: int = -1
    sensor_visual_rab_ring_segments: int = 48

    corridor_width: float = 0.50
    corridor_length: float = 1.06
    gate_width: float = 0.45
    gate_length: float = 0.33
    side_wall_length: float = 0.50

    light_position: tuple = (0.0, -1.5, 0.0)
    has_light: bool = True
    alpha_parameter: float = 5.0

    # mission key consumed by params.build_params; not a reference attribute
    _mission: str = "dgt"

    def update_variant(self, variant: str):
        """CFG:184-193."""
        if variant not in OBS_DIM:
            raise KeyError(variant)
        self.variant = variant
        self.observation_spaces = _spaces(OBS_DIM[variant], self.num_agents)
        self.action_spaces = _spaces(ACT_DIM[variant], self.num_agents)
        self.discrete_actions = variant != "dandelion"

    def use_continuous_actions(self, full_observations: bool = False):
        """CFG:195-209."""
        self.action_spaces = _spaces(2, self.num_agents)
        self.discrete_actions = False
        self.full_policy_observations = bool(full_observations)
        if self.full_policy_observations:
            self.observation_spaces = _spaces(24, self.num_agents)

    def copy(self):
        return copy.deepcopy(self)

    @property
    def obs_dim(self) -> int:
        """ENV:1132-1141: 24-dim iff dandelion/daisy or full_policy_observations."""
        return 24 if (self.variant in ("dandelion", "daisy") or self.full_policy_observations) else 4


@dataclass
class XorAggregationEnvCfg(DirectionalGateEnvCfg):
    episode_length_s: float = 180.0
    has_light: bool = False
    spawn_area_size: tuple = (2.4, 2.4)
    spawn_circle_radius: float = 1.2
    target_radius: float = 0.30
    target_centers: tuple = ((-0.50, 0.0), (0.50, 0.0))
    _mission: str = "xor"


@dataclass
class HomingEnvCfg(DirectionalGateEnvCfg):
    episode_length_s: float = 120.0
    has_light: bool = False
    spawn_area_center: tuple = (0.0, 0.7)
    spawn_area_size: tuple = (2.0, 0.6)
    spawn_circle_radius: float = 0.8
    goal_radius: float = 0.30
    goal_center: tuple = (0.0, -0.70)
    _mission: str = "hom"


@dataclass
class ForagingEnvCfg(DirectionalGateEnvCfg):
    episode_length_s: float = 180.0
    has_light: bool = True
    light_position: tuple = (0.0, -1.5, 0.0)
    spawn_area_size: tuple = (1.8, 1.8)
    spawn_circle_radius: float = 0.0
    food_radius: float = 0.15
    food_centers: tuple = ((-0.75, 0.0), (0.75, 0.0))
    nest_top_y: float = -0.58
    _mission: str = "for"


@dataclass
class ShelteringEnvCfg(DirectionalGateEnvCfg):
    episode_length_s: float = 180.0
    has_light: bool = True
    light_position: tuple = (0.0, -1.5, 0.0)
    spawn_area_size: tuple = (1.8, 1.8)
    spawn_circle_radius: float = 0.0
    shelter_center: tuple = (0.0, 0.0)
    shelter_size: tuple = (0.50, 0.30)
    shelter_wall_thickness: float = 0.03
    black_area_radius: float = 0.30
    black_area_centers: tuple = ((-0.80, 0.0), (0.80, 0.0))
    _mission: str = "shl"


# Gymnasium ids of the reference (missions/*/__init__.py) -> cfg class
TASK_CFGS = {
    "SwarmACB-DirectionalGate-v0": DirectionalGateEnvCfg,
    "SwarmACB-XOR-v0": XorAggregationEnvCfg,
    "SwarmACB-Homing-v0": HomingEnvCfg,
    "SwarmACB-Foraging-v0": ForagingEnvCfg,
    "SwarmACB-Sheltering-v0": ShelteringEnvCfg,
    "SwarmACB-SCA-v0": ShelteringEnvCfg,
    "SwarmACB-SHL-v0": ShelteringEnvCfg,
}

MISSION_CFGS = {
    "dgt": DirectionalGateEnvCfg, "xor": XorAggregationEnvCfg, "hom": HomingEnvCfg,
    "for": ForagingEnvCfg, "shl": ShelteringEnvCfg,
}
