"""trimPathNearOccupiedRegions (src/aos_path_gen_node.cpp:1570-1630): aos_trim_path on the
device against the oracle's restatement of the reference loop -- same pose count, on random paths over random and
orchard skeletons, at the grid border, with NaN poses, with a pose 0 that is too close (never trims)."""
import numpy as np
import pytest

from aos_gpu import lib, synth
from helpers import params_pair

pytestmark = pytest.mark.gpu


def _random_grid(rng, h, w, density):
    g = np.zeros((h, w), np.int8)
    g[rng.random((h, w)) < density] = 100
    g[0, :] = g[-1, :] = 100
    g[:, 0] = g[:, -1] = 100          # the published skeleton carries the 1-px frame
    return g


def test_trim_matches_reference_loop_random(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    for trial in range(60):
        h, w = int(rng.integers(5, 140)), int(rng.integers(5, 170))
        res = [0.05, 0.1, 0.025, 0.07][trial % 4]
        ox, oy = [(0.0, 0.0), (-4.5, -2.4), (1000.25, 37.125)][trial % 3]
        g = _random_grid(rng, h, w, [0.0005, 0.003, 0.02][trial % 3])
        n = int(rng.integers(0, 400))
        # a random walk that starts in the interior and wanders over (and sometimes off) the grid
        start = np.array([ox + w * res * rng.uniform(0.2, 0.8), oy + h * res * rng.uniform(0.2, 0.8)])
        path = start + np.cumsum(rng.normal(0, res * 1.5, (n, 2)), axis=0)
        if n > 5 and trial % 7 == 0:
            path[3] = [np.nan, path[3, 1]]
        if n > 5 and trial % 11 == 0:   # pose 0 on an occupied cell: tested, never trims
            yy, xx = np.argwhere(g == 100)[len(g) // 2 % max(1, (g == 100).sum())]
            path[0] = [ox + (xx + 0.5) * res, oy + (yy + 0.5) * res]
        safety = [0.2, 0.2, 0.35, 0.0][trial % 4]
        want = oracle.trim_path(path, g, ox, oy, res, safety)
        got = gpu_ctx.trim_path(path, safety, skeleton_bits=lib.pack_bits(g == 100), info=(res, ox, oy, w))
        assert got == want, f"trial {trial}: kept {got} vs {want} of {n}"


def test_trim_on_context_skeleton(gpu_ctx, oracle):
    """The skeleton of this context's seed stage (no grid handed over): paths along the alleys survive, a path that
    crosses a tree row is cut in front of it."""
    spec = synth.config("SMALL", seed=4)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    res, ox, oy, w, h = float(r["res"]), r["origin_x"], r["origin_y"], r["w"], r["h"]
    rng = np.random.default_rng(2)
    cut = 0
    for trial in range(30):
        a = np.array([ox + rng.uniform(1, w * res - 1), oy + rng.uniform(1, h * res - 1)])
        b = np.array([ox + rng.uniform(1, w * res - 1), oy + rng.uniform(1, h * res - 1)])
        t = np.linspace(0, 1, int(rng.integers(2, 800)))[:, None]
        path = a + (b - a) * t
        want = oracle.trim_path(path, r["skel_framed"], ox, oy, res, 0.2)
        got = gpu_ctx.trim_path(path)
        assert got == want
        cut += want < len(path)
    assert cut > 0          # some of the random chords do cross a row


def test_trim_edge_cases(gpu_ctx):
    g = np.zeros((40, 40), np.int8)
    g[20, 20] = 100
    bits = lib.pack_bits(g == 100)
    info = (0.05, 0.0, 0.0, 40)
    assert gpu_ctx.trim_path(np.zeros((0, 2)), 0.2, skeleton_bits=bits, info=info) == 0
    assert gpu_ctx.trim_path(np.array([[1.0, 1.0]]), 0.2, skeleton_bits=bits, info=info) == 1     # pose 0 only
    path = np.stack([np.linspace(0.1, 1.9, 50), np.full(50, 1.025)], 1)
    k = gpu_ctx.trim_path(path, 0.2, skeleton_bits=bits, info=info)
    assert 0 < k < 50 and path[k, 0] >= 1.0 - 0.2 - 0.05 - 1e-9 and path[k - 1, 0] < path[k, 0]
    # no skeleton on a fresh context: AOS_ERR_STATE (the reference leaves the path alone, path_gen:1571)
    c = lib.Context(0)
    with pytest.raises(lib.AosError):
        c.trim_path(path)
    c.close()
