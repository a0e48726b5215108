"""BASELINE config 1: the oracle's manual-control tick against the reference's StandaloneDGTEnv rollouts."""
import pytest

import fixtures
from oracle import oracle

FILES = fixtures.mc_fixture_files()


def test_mc_fixtures_present():
    assert len(FILES) == 5


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_manual_control_rollout(path):
    fx = fixtures.McFixture(path)
    for t in range(fx.ticks):
        state = fx.state(t)
        obs, reward, rolled = oracle.mc_tick(fx.params, state, **fx.tick_inputs(t))
        fx.check(t, state, obs, reward, rolled, label=f"oracle:{fx.name}")
