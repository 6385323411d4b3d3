"""Auction restatement (hungarian.cu:27-123, 358-405): matching properties, the stated
tie-breaks, equality of the early-stopped loop with the literal fixed-iteration loop."""
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_partial_matching(row, col):
    for r, c in enumerate(row):
        if c >= 0:
            assert col[c] == r
    for c, r in enumerate(col):
        if r >= 0:
            assert row[r] == c


@pytest.mark.parametrize("R,C,seed", [(5, 5, 0), (20, 20, 1), (50, 50, 2), (30, 12, 3), (12, 30, 4), (128, 64, 5)])
def test_matching_is_consistent(orc, R, C, seed):
    rng = np.random.default_rng(seed)
    cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
    row, col = orc.auction(cost)
    check_partial_matching(row, col)
    assert (row >= 0).sum() == (col >= 0).sum() <= min(R, C)


def test_small_square_problems_are_near_optimal(orc):
    """With eps <= 1/(n+1) decreasing, a converged auction is within n*eps of optimal."""
    rng = np.random.default_rng(7)
    for _ in range(20):
        n = 5
        cost = rng.uniform(0, 1, (n, n)).astype(np.float32)
        row, col = orc.auction(cost)
        if (row >= 0).sum() < n:
            continue
        got = sum(cost[r, row[r]] for r in range(n))
        best = min(sum(cost[r, p[r]] for r in range(n)) for p in itertools.permutations(range(n)))
        assert got <= best + n * (1.0 / (n + 1)) + 1e-6


def test_inactive_rows_never_assigned(orc):
    rng = np.random.default_rng(9)
    cost = rng.uniform(0, 1, (16, 8)).astype(np.float32)
    active = (np.arange(16) % 3 == 0).astype(np.int32)
    row, col = orc.auction(cost, active)
    assert (row[active == 0] == -1).all()
    check_partial_matching(row, col)


def test_locked_cells_are_never_chosen(orc):
    """cost 1e9 gives value -1e9 - price, which is not > -1e9 (hungarian.cu:55-63)."""
    cost = np.full((4, 4), 1e9, np.float32)
    cost[1, 2] = 0.25
    row, col = orc.auction(cost)
    assert list(row) == [-1, 2, -1, -1] and list(col) == [-1, -1, 1, -1]


def test_tie_breaks(orc):
    # equal values: the lowest column is bid on; equal bids: the lowest row wins (R6)
    cost = np.full((3, 3), 0.5, np.float32)
    row, col = orc.auction(cost)
    check_partial_matching(row, col)
    cost = np.array([[0.1, 1e9], [0.1, 1e9]], np.float32)   # both rows can only take column 0
    row, col = orc.auction(cost)
    assert col[0] == 0 and row[0] == 0 and row[1] == -1


def test_early_stop_equals_literal_iterations():
    """ORC_LITERAL_AUCTION=1 runs all min(3R,50) iterations; results must be identical."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_py as orc
rng = np.random.default_rng(11)
out = []
for R, C in [(8, 8), (40, 25), (25, 40), (128, 64), (64, 20)]:
    cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
    cost[rng.uniform(size=(R, C)) < 0.3] = 1e9
    act = (rng.uniform(size=R) < 0.7).astype(np.int32)
    row, col = orc.auction(cost, act)
    out.append(row.tolist()); out.append(col.tolist())
print(out)
""" % (ROOT, os.path.join(ROOT, "oracle"))
    outs = []
    for lit in ("0", "1"):
        env = dict(os.environ, ORC_LITERAL_AUCTION=lit)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1] and len(outs[0]) > 100
