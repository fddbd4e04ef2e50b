"""Work decomposition of the streamed passes (pure host logic, no device needed).

A pass over P rows x Q columns is cut into work items (row tile, column split).  The SIMT kernels use 64-row tiles and
64-column streamed tiles, the tensor-core kernel 128-row tiles and 256-column tiles.  Every plan keeps a split at or
below `max_split_cols` columns: the kernels accumulate a split's sum in fp32 and 65536 terms keep that below 1e-6
relative (include/spadot_b200.h).
"""
from __future__ import annotations


def aligned_bounds(n_q: int, ns: int, tile: int = 64):
    """ns + 1 column boundaries of n_q columns, every inner boundary on a `tile` multiple (no item streams a mostly
    masked tile).  ns must not exceed the number of tiles."""
    tiles = max(1, -(-n_q // tile))
    return [min(n_q, tile * ((tiles * s) // ns)) for s in range(ns)] + [n_q]


def simt_splits(n_p: int, n_q: int, target_ctas: int, max_split_cols: int) -> int:
    """Number of column splits of the one-item-per-CTA SIMT pass: about `target_ctas` CTAs in total, at least one
    64-column tile per item for small problems and four for large ones, never more columns per split than the cap."""
    row_tiles = max(1, (n_p + 63) // 64)
    want = max(1, -(-target_ctas // row_tiles))
    lo = max(1, -(-n_q // max_split_cols))
    tiles = max(1, -(-n_q // 64))
    hi = max(1, n_q // 256) if n_q >= 4096 else tiles
    return int(min(max(want, lo), max(hi, lo), 65535))


def persistent_splits(n_p: int, n_q: int, grid: int, max_split_cols: int) -> int:
    """Column splits of the persistent small-problem kernel: about one item per CTA of the cooperative grid."""
    row_tiles = max(1, (n_p + 63) // 64)
    tiles = max(1, -(-n_q // 64))
    return max(min(max(1, grid // row_tiles), tiles), -(-n_q // max_split_cols))


def tc_split_plan(n_p: int, n_q: int, n_ctas: int, max_split_cols: int, item_overhead_tiles: float):
    """(tiles per split, number of splits) of one tensor-core pass.  Work items (128-row tile, split) are dealt
    round-robin to the persistent CTAs, so the pass lasts  ceil(items / CTAs) * (tiles per item + per-item overhead):
    pick the split length that minimises it (ties: longer items)."""
    row_tiles = (n_p + 127) // 128
    col_tiles = (n_q + 255) // 256
    best = None
    for tps in range(1, min(col_tiles, max_split_cols // 256) + 1):
        ns = -(-col_tiles // tps)
        rounds = -(-(row_tiles * ns) // n_ctas)
        cost = rounds * (tps + item_overhead_tiles)
        if best is None or cost <= best[0]:
            best = (cost, tps, ns)
    return best[1], best[2]
