import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _forced_tile_schedule():
    """BG_TEST_TILE_SCHEDULE=0|1 runs the whole GPU suite with the tcgen05 evaluator's static / dynamic tile schedule forced for every
    launch (the default picks by size), so that both code paths see every test."""
    mode = os.environ.get("BG_TEST_TILE_SCHEDULE")
    if mode in ("0", "1"):
        import mlp_ppo_2ply_multi_b200 as bg

        bg._lib.lib().bg_eval_tc_tile_schedule(int(mode))
    yield


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/bg_oracle.c) -- the checker, never the thing under test on the GPU path."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden
