"""Pin the CPU oracle against the golden fixtures generated from the unmodified reference."""
import copy

import numpy as np
import pytest

import fixtures
from oracle import oracle

FILES = fixtures.fixture_files()


def test_fixtures_present():
    assert len(FILES) >= 60


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_step(path):
    fx = fixtures.Fixture(path)
    for t in range(fx.steps):
        case = fx.step_case(t)
        state = copy.deepcopy(case["pre"])
        obs, reward, time_out = oracle.step(
            fx.params, state, case["actions"], rab_u=case["rab_u"], turn_dur=case["turn_dur"],
            spawn_u=case["spawn_u"], yaw_u=case["yaw_u"])
        critic = oracle.critic_state(fx.params, state)
        fixtures.compare(case, fx.params, state, obs, reward, time_out, critic, label=f"{fx.name}[{t}]")
