import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Tolerance of BASELINE.json's north_star: distances within 1e-5 relative; near a
# self-match the distance itself is ~0 and 1 - cos cancels, so the bound is taken
# as |delta| <= 1e-5 * max(|d|, 1) (SURVEY.md §8c "Tolerance note").
RTOL = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def tol(d):
    return RTOL * np.maximum(np.abs(np.asarray(d, dtype=np.float64)), 1.0)


def assert_topk_parity(got_pos, got_dist, exp_pos, exp_dist, dist_all):
    """GPU result vs oracle result for one query.

    ``*_pos`` are scan positions, ``dist_all`` the oracle's distance for every
    position.  Row ids must be identical except where the oracle's distances of
    the two rows tie within the tolerance; every returned distance must be
    within tolerance of the oracle's distance for that same row.
    """
    got_pos = np.asarray(got_pos)
    exp_pos = np.asarray(exp_pos)
    got_dist = np.asarray(got_dist, dtype=np.float64)
    exp_dist = np.asarray(exp_dist, dtype=np.float64)
    assert got_pos.shape == exp_pos.shape, (got_pos.shape, exp_pos.shape)
    assert len(set(got_pos.tolist())) == len(got_pos), "duplicate rows in result"
    own = np.asarray(dist_all, dtype=np.float64)[got_pos]
    assert np.all(np.abs(got_dist - own) <= tol(own)), "distance off for the returned row"
    assert np.all(np.abs(got_dist - exp_dist) <= tol(exp_dist)), "rank-wise distance off"
    assert np.all(np.diff(got_dist) >= 0), "result not sorted by distance"
    diff = got_pos != exp_pos
    if diff.any():
        # only near-ties may swap
        assert np.all(np.abs(own[diff] - exp_dist[diff]) <= tol(exp_dist[diff])), \
            f"row ids differ beyond a tolerance tie at ranks {np.nonzero(diff)[0].tolist()}"
    return int(diff.sum())


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_search.json")) as f:
        return json.load(f)
