"""Host-side work decomposition (spadot_b200/plans.py): every plan covers all columns exactly once, respects the
fp32 accumulation cap per split and never produces an empty item."""
import itertools

import pytest

from spadot_b200 import plans

CAP = 65536
SIZES = [1, 5, 63, 64, 65, 300, 747, 1966, 4096, 5000, 18408, 100000, 1_000_000]


@pytest.mark.parametrize("n_p,n_q", list(itertools.product([1, 747, 2000, 50000, 1_000_000], SIZES)))
def test_simt_and_persistent_splits_cover_columns(n_p, n_q):
    for ns in (plans.simt_splits(n_p, n_q, 148 * 6, CAP), plans.persistent_splits(n_p, n_q, 296, CAP)):
        b = plans.aligned_bounds(n_q, ns)
        assert len(b) == ns + 1 and b[0] == 0 and b[-1] == n_q
        assert all(lo < hi for lo, hi in zip(b[:-1], b[1:])), (n_p, n_q, ns, b[:5])          # no empty split
        assert all(hi - lo <= CAP + 64 for lo, hi in zip(b[:-1], b[1:]))                     # accumulation cap (+ one tile)
        assert all(x % 64 == 0 for x in b[1:-1])                                             # inner cuts on tile boundaries
        assert 1 <= ns <= 65535


@pytest.mark.parametrize("n_p,n_q", list(itertools.product([1, 128, 4096, 8192, 125000, 1_000_000], [1, 256, 4100, 8192, 32768, 1_000_000])))
@pytest.mark.parametrize("over", [0.5, 2.0])
def test_tc_split_plan(n_p, n_q, over):
    tps, ns = plans.tc_split_plan(n_p, n_q, 148, CAP, over)
    col_tiles = (n_q + 255) // 256
    assert 1 <= tps <= CAP // 256
    assert ns == -(-col_tiles // tps)
    assert (ns - 1) * tps < col_tiles <= ns * tps                                            # the last split is not empty
    # the chosen plan is no worse than the two obvious ones under the planner's own cost model
    row_tiles = (n_p + 127) // 128

    def cost(t):
        return -(-(row_tiles * -(-col_tiles // t)) // 148) * (t + over)
    assert cost(tps) <= cost(min(col_tiles, CAP // 256)) and cost(tps) <= cost(1)


def test_tc_plan_examples_from_the_design_notes():
    # 8192^2: 64 row tiles x 32 column tiles on 148 CTAs -> one round of 16-tile items beats two rounds of 8-tile items
    assert plans.tc_split_plan(8192, 8192, 148, CAP, 2.0) == (16, 2)
    # 1M x 1M: splits stay under the 65536-column cap
    tps, ns = plans.tc_split_plan(1_000_000, 1_000_000, 148, CAP, 2.0)
    assert tps <= 256 and ns >= 16
