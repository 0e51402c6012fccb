import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUT = os.path.join(GOLDEN, "input")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "ref_out_121144_com2.npz")))


@pytest.fixture(scope="session")
def case121144():
    """The reference's own case: feeder 121144, community 2, 90 % adoption, 4800 W
    (test-dist-ind-opt.py / revs_config.yaml), read through this package's extract.py."""
    from revs_admm_b200.revs_fixture import REVS
    fx = REVS(data_path=INPUT, out_path="/tmp/revs_out", grb_path="/tmp/revs_grb",
              fig_path="/tmp/revs_fig", regionID=121, networkID=121144, comunityID=2,
              optimizer_mode="distributed")
    tariff, homes, dist, saved = fx.read_inputs(adoption=90, rating=4800, seed=1234, capacity=20,
                                                initial_soc=0.2, start_time=11, end_time=23,
                                                shift_time=6)
    return dict(fx=fx, tariff=tariff, homes=homes, dist=dist, saved=saved)


@pytest.fixture(scope="session")
def gpu_lib():
    import revs_admm_b200 as r
    if r.device_count() < 1:
        pytest.fail("no CUDA device visible")
    return r


def _case(adoption, rating):
    from revs_admm_b200.revs_fixture import REVS
    fx = REVS(data_path=INPUT, out_path="/tmp/revs_out", grb_path="/tmp/revs_grb",
              fig_path="/tmp/revs_fig", regionID=121, networkID=121144, comunityID=2,
              optimizer_mode="individual")
    tariff, homes, dist, saved = fx.read_inputs(adoption=adoption, rating=rating, seed=1234, capacity=20,
                                                initial_soc=0.2, start_time=11, end_time=23,
                                                shift_time=6)
    return dict(fx=fx, tariff=tariff, homes=homes, dist=dist, saved=saved)


@pytest.fixture(scope="session")
def case_adopt70():
    """Second reference run shipped in out/121144-com2/individual: 70 % adoption, 4800 W."""
    return _case(70, 4800)


@pytest.fixture(scope="session")
def case_rating3600():
    """Third reference run shipped in out/121144-com2/individual: 90 % adoption, 3600 W."""
    return _case(90, 3600)
