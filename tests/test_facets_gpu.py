"""cv::Subdiv2D::calcVoronoi + getVoronoiFacetList on the device (csrc/k_facets.cu) against the real cv2.Subdiv2D
(through oracle/subdiv.py) and against the host walk of the same library, bit for bit: the facet-vertex slots the
graph kernels consume (facets with fewer than 2 vertices dropped, voronoi_diagram.cpp:97-114) and their cycle links."""
import numpy as np
import pytest

from aos_gpu import lib
from test_subdiv_cpu import _seed_sets

pytestmark = pytest.mark.gpu


def _slots_from_facets(xy, off):
    """What gvd_stage builds from a facet list: one slot per vertex of every facet with >= 2 vertices."""
    keep_xy, nxt, base = [], [], 0
    for f in range(len(off) - 1):
        b, k = int(off[f]), int(off[f + 1] - off[f])
        if k < 2:
            continue
        keep_xy.append(xy[b:b + k])
        n = np.arange(base + 1, base + k + 1, dtype=np.int32)
        n[-1] = base
        nxt.append(n)
        base += k
    if not keep_xy:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.int32)
    return np.concatenate(keep_xy), np.concatenate(nxt)


@pytest.fixture(scope="module")
def ctx():
    c = lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("block", range(8))
def test_device_facets_match_cv2_bit_for_bit(oracle, ctx, block):
    from oracle import subdiv
    rng = np.random.default_rng(100 + block)
    for trial in range(40):
        s = _seed_sets(rng, trial)
        b = (0.0, 50.0, 0.0, 40.0) if trial % 3 else (-4.5, 72.8, -2.4, 12.4)
        fx, fo, _ = subdiv.voronoi_facets(s, *b)
        want_xy, want_next = _slots_from_facets(fx, fo)
        got_xy, got_next = ctx.voronoi_facets_device(s, *b)
        assert got_xy.shape == want_xy.shape, f"slot count differs (block {block} trial {trial})"
        assert np.array_equal(got_next, want_next), f"cycle links differ (block {block} trial {trial})"
        assert np.array_equal(got_xy.view(np.uint32), want_xy.view(np.uint32)), f"vertices differ (block {block} trial {trial})"


def test_device_facets_large_orchard_equals_host_walk(ctx):
    """60 k seeds in rows (the insertion pattern of a real map; far beyond what cv2 is asked for in the small tests):
    device walk == host walk of the same subdivision; second call re-uses the pinned arrays."""
    rng = np.random.default_rng(5)
    rows = []
    for r in range(120):
        x = np.arange(500) * 1.0 + rng.uniform(0, 1)
        rows.append(np.stack([x + rng.normal(0, 0.02, 500), 3 + 4 * r + 0.01 * x * rng.uniform(-1, 1)], 1))
    s = np.concatenate(rows)
    b = (0.0, 520.0, 0.0, 500.0)
    hx, ho = lib.voronoi_facets(s, *b)
    want_xy, want_next = _slots_from_facets(hx, ho)
    for _ in range(2):
        got_xy, got_next = ctx.voronoi_facets_device(s, *b)
        assert np.array_equal(got_next, want_next)
        assert np.array_equal(got_xy.view(np.uint32), want_xy.view(np.uint32))
    # a smaller set afterwards on the same context (arrays shrink logically, stay pinned)
    got_xy, got_next = ctx.voronoi_facets_device(s[:777], *b)
    hx, ho = lib.voronoi_facets(s[:777], *b)
    want_xy, want_next = _slots_from_facets(hx, ho)
    assert np.array_equal(got_next, want_next) and np.array_equal(got_xy.view(np.uint32), want_xy.view(np.uint32))


def test_device_facets_empty(ctx):
    xy, nxt = ctx.voronoi_facets_device(np.zeros((0, 2)), 0, 10, 0, 10)
    assert len(xy) == 0 and len(nxt) == 0
    xy, nxt = ctx.voronoi_facets_device(np.array([[1.0, 2.0]]), 0, 10, 0, 10)   # one seed: one facet of the 3 outer circumcentres
    hx, ho = lib.voronoi_facets(np.array([[1.0, 2.0]]), 0, 10, 0, 10)
    want_xy, want_next = _slots_from_facets(hx, ho)
    assert np.array_equal(nxt, want_next) and np.array_equal(xy.view(np.uint32), want_xy.view(np.uint32))


def test_device_merge_matches_reference_loop(oracle, ctx):
    """voronoiSeedsCallback's greedy 0.5 m merge on the device vs the oracle's literal O(S^2) loop (gvd:84-128)."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.integers(0, 900))
        s = rng.uniform(0, 12 if trial % 2 else 6, (n, 2))
        if n > 10:
            s[5] = s[4] + [0.5, 0.0]      # exactly at the merge distance (<=)
            s[9] = [np.nan, 1.0]
            s[20:28] = s[19]              # a pile of identical seeds
        if trial % 5 == 0 and n > 50:     # chains: every seed within 0.5 m of its predecessor
            s[30:50, 0] = 3 + 0.3 * np.arange(20)
            s[30:50, 1] = 3
        want = oracle.merge_seeds(s)
        want = want[np.isfinite(want).all(axis=1)]   # processGraph drops non-finite merged seeds (gvd:266-270)
        got = ctx.merge_seeds_device(s)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"trial {trial}"
        assert np.array_equal(got, lib.merge_seeds(s))


def test_device_merge_dense_pile(ctx):
    """2000 seeds inside one 0.5 m disc plus a far one: one leader absorbs them all, in index order."""
    rng = np.random.default_rng(3)
    s = np.concatenate([rng.uniform(0, 0.2, (2000, 2)) + 5.0, [[50.0, 50.0]]])
    got = ctx.merge_seeds_device(s)
    want = lib.merge_seeds(s)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
