"""GPU parity of the seed-gen half against the oracle, through the C-ABI (bit-exact)."""
import numpy as np
import pytest

from aos_gpu import lib, synth
from helpers import assert_seed_parity, params_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,seed", [("TINY", 0), ("TINY", 1), ("SMALL", 0), ("SMALL", 3)])
def test_seed_stage_small(gpu_ctx, oracle, name, seed):
    spec = synth.config(name, seed=seed)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)


def test_seed_stage_c1(gpu_ctx, oracle):
    spec = synth.config("C1", n_points=400_000)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)


def test_seed_stage_c2_device_points(gpu_ctx, oracle):
    import torch
    spec = synth.config("C2", n_points=600_000)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    dpts = torch.from_numpy(pts).cuda()
    gpu_ctx.seed_stage(pl, dpts)
    assert_seed_parity(gpu_ctx, r)


def test_reference_polygon_and_exclusion_discs(gpu_ctx, oracle):
    """The reference's own field constants: default polygon (seed_gen:196-199, negative origin,
    non-rectangular) and the 11 exclusion discs (seed_gen:487-499)."""
    spec = synth.OrchardSpec(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5,
                             n_points=300_000, seed=5, exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle, polygon=synth.REFERENCE_POLYGON)
    r = oracle.seed_stage(po, pts)
    assert (r["w"], r["h"]) == (1546, 296)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)


def test_generic_point_step(gpu_ctx, oracle):
    """32-byte XYZI-style records with the fields at odd offsets (PointCloud2 point_step path)."""
    spec = synth.config("TINY", seed=2)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    rec = np.zeros((len(pts), 8), np.float32)
    rec[:, 1], rec[:, 2], rec[:, 5] = pts[:, 0], pts[:, 1], pts[:, 2]
    gpu_ctx.seed_stage(pl, rec, point_step=32, offsets=(4, 8, 20))
    assert_seed_parity(gpu_ctx, r)


def test_empty_and_no_polygon(gpu_ctx, oracle):
    po = oracle.SeedParams(clipping_minx=-1.0, clipping_maxx=9.0, clipping_miny=-2.0, clipping_maxy=5.0)
    pl = lib.SeedParams(clipping_minx=-1.0, clipping_maxx=9.0, clipping_miny=-2.0, clipping_maxy=5.0)
    pts = np.zeros((0, 4), np.float32)
    r = oracle.seed_stage(po, np.zeros((1, 4), np.float32) + np.float32(np.nan))
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)
    rng = np.random.default_rng(0)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, 0] = rng.uniform(-2, 10, 5000)
    pts[:, 1] = rng.uniform(-3, 6, 5000)
    pts[:, 2] = rng.uniform(-1, 1, 5000)
    pts[::97, 0] = np.nan
    pts[5::101, 2] = np.inf
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)


def test_long_rows_bfs_order_replay(gpu_ctx, oracle):
    """1 km rows: float32 running sums pass 2^24, so the centre depends on the reference's BFS order
    (seed_gen:1053-1059); the library replays that order (bfs_replay_kernel)."""
    spec = synth.OrchardSpec(extent_x=1000.0, extent_y=16.0, row_pitch=4.0, n_points=500_000, gap_prob=0.0,
                             jitter=0.0, outlier_count=4, seed=11)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    assert r["cl_sumx"].max() >= (1 << 24)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r, check_labels=False)


@pytest.mark.parametrize("case", ["long_rows", "tie_breaks", "rotated"])
def test_literal_bfs_replay_equals_chain_walk(gpu_ctx, oracle, case, monkeypatch):
    """The BFS order comes from the chain-compressed walk (bfs_chain_kernel); the literal warp-per-cluster replay
    (bfs_replay_kernel) is its fallback for levels wider than the walk's list.  AOS_LITERAL_BFS forces the fallback for
    every flagged cluster: both must reproduce the oracle (seed_gen:1008-1059), on rows, fragments and rotated rows."""
    if case == "long_rows":
        spec = synth.OrchardSpec(extent_x=1000.0, extent_y=16.0, row_pitch=4.0, n_points=500_000, gap_prob=0.0,
                                 jitter=0.0, outlier_count=4, seed=11)
        over = {}
    elif case == "tie_breaks":
        spec = synth.config("TINY", seed=1)
        spec.outlier_count = 12
        over = dict(cluster_min_length=0.0)
    else:
        spec = synth.OrchardSpec(extent_x=400.0, extent_y=100.0, row_pitch=5.0, n_points=800_000, gap_prob=0.01,
                                 jitter=0.05, outlier_count=8, seed=5, rotation_deg=17.0)
        over = {}
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle, **over)
    r = oracle.seed_stage(po, pts)
    if case != "tie_breaks":
        assert max(r["cl_sumx"].max(), r["cl_sumy"].max()) >= (1 << 24)  # some centre depends on the order
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r, check_labels=False)
    monkeypatch.setenv("AOS_LITERAL_BFS", "1")
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r, check_labels=False)
    # a level list of two cells: clusters with a junction hand over to the fallback, plain rows stay with the walk
    monkeypatch.delenv("AOS_LITERAL_BFS")
    monkeypatch.setenv("AOS_BFS_ITEM_CAP", "2")
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r, check_labels=False)
    # the walk on the global tables (what a cluster too large for the shared-memory staging does)
    monkeypatch.delenv("AOS_BFS_ITEM_CAP")
    monkeypatch.setenv("AOS_BFS_GLOBAL", "1")
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r, check_labels=False)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_min_length_zero_tie_breaks(gpu_ctx, oracle, seed):
    """cluster_min_length = 0 turns every fragment into a row, including tiny symmetric ones whose
    farthest-cell arg-max ties are broken by BFS order in the reference (seed_gen:1359-1367)."""
    spec = synth.config("TINY", seed=seed)
    spec.outlier_count = 12
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle, cluster_min_length=0.0)
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    assert_seed_parity(gpu_ctx, r)
