"""Full-size checks (BASELINE.json config: 125k homes x 96 steps per GPU) through size-independent
properties of the path -- the CPU oracle cannot run this size in seconds.  See
profiles/check_properties.py for what is checked."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("check_properties", os.path.join(ROOT, "profiles", "check_properties.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", [0, 5])
def test_projection_is_feasible_at_full_size(gpu_lib, seed):
    cp = _tool()
    trees, pe, gm, res, st = cp.run("synthetic-multifeeder-125k-homes-per-gpu-x96", seed)
    assert pe.min() >= 0.0
    assert cp.feasibility(trees, pe, zones=120) <= 1e-9           # R P_est <= u in every sampled zone
    assert st["max_working_set"] >= 8                              # the limits bind
    # schedule bookkeeping: P_sch = load + P_ev, P_ev in {0, rating}, SOC recursion and final SOC window
    assert np.all((res["P_ev"] == 0.0) | (res["P_ev"] > 0.0))
    soc_end = res["SOC"][:, -1]
    ev = res["P_ev"].max(axis=1) > 0
    assert np.all(soc_end[ev] >= 0.9 - 1e-9) and np.all(soc_end <= 1.0 + 1e-9)


def test_modes_agree_bitwise_at_full_size(gpu_lib):
    cp = _tool()
    wl = "synthetic-multifeeder-125k-homes-per-gpu-x96"
    _, pe, gm, res, _ = cp.run(wl, 0)
    _, pe2, gm2, res2, _ = cp.run(wl, 0, stepwise=True)
    _, pe3, gm3, res3, _ = cp.run(wl, 0, screen=0)
    for a, b in ((pe, pe2), (gm, gm2), (res["P_sch"], res2["P_sch"]), (pe, pe3), (gm, gm3), (res["P_sch"], res3["P_sch"])):
        assert np.array_equal(a, b)
