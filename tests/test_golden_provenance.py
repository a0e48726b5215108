"""The committed fixtures ARE the reference's outputs: where the reference checkout is present (the build
container), a sample of tests/golden/*.npz is regenerated from the unmodified reference classes with the committed
generator and compared array by array.  Skipped on the GPU box (no reference there)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present")


def _cases():
    sys.path.insert(0, GOLD)
    import gen_golden
    return gen_golden, gen_golden.case_list()


@pytest.mark.parametrize("pick", [0, 7, 13, 25, 38, 52, 64])
def test_fixture_regenerates_from_the_live_reference(pick):
    gen_golden, cases = _cases()
    mission, mode, scenario, E, steps, dec, seed = cases[pick % len(cases)]
    meta, arrays = gen_golden.make_case(mission, mode, scenario, E, steps, seed, decimation=dec)
    path = os.path.join(GOLD, f"{mission}_{mode}_{scenario}_d{dec}.npz")
    with np.load(path) as z:
        stored_meta = json.loads(str(z["meta"]))
        assert {k: v for k, v in stored_meta.items() if k != "torch"} == {k: v for k, v in meta.items() if k != "torch"}
        assert set(z.files) - {"meta"} == set(arrays)
        for k, v in arrays.items():
            a = z[k]
            assert a.shape == v.shape and a.dtype == v.dtype, k
            if a.dtype.kind == "f":   # same torch build -> identical bits; a different build may differ in the last ulp
                assert np.allclose(a, v, rtol=0, atol=2e-6, equal_nan=True), f"{path}: {k}"
            else:
                assert np.array_equal(a, v), f"{path}: {k}"
