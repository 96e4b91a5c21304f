"""aos_radius_outlier_removal (row F1, pcl::RadiusOutlierRemoval ahead of the seam, seed_gen:229-248) against the
brute-force restatement in the oracle.  PCL itself is not available: parity unpinned against the real thing."""
import numpy as np
import pytest

from aos_gpu import lib, synth
from helpers import assert_seed_parity, params_pair

pytestmark = pytest.mark.gpu


def _fetch(dp):
    import torch
    from aos_gpu.bands import _CudaArray
    if dp.shape[0] == 0:
        return np.zeros((0, 4), np.float32)
    t = torch.as_tensor(_CudaArray(dp.ptr, (dp.shape[0], 4), "<f4"), device="cuda")
    return t.cpu().numpy()


def _cloud(rng, n_dense, n_sparse):
    centres = rng.uniform(0, 20, (12, 3)) * [1, 1, 0.1]
    d = centres[rng.integers(0, 12, n_dense)] + rng.normal(0, 0.25, (n_dense, 3))
    s = rng.uniform(-2, 22, (n_sparse, 3)) * [1, 1, 0.2]
    pts = np.concatenate([d, s]).astype(np.float32)
    rng.shuffle(pts, axis=0)
    return pts


@pytest.mark.parametrize("seed,radius,k", [(0, 0.2, 2), (1, 0.2, 2), (2, 0.35, 5), (3, 0.1, 1), (4, 0.2, 0)])
def test_ror_matches_restatement(gpu_ctx, oracle, seed, radius, k):
    rng = np.random.default_rng(seed)
    pts = _cloud(rng, 4000, 1500)
    pts[7] = pts[3]                       # exact duplicates count as neighbours
    pts[11] = pts[3]
    pts[20, 1] = np.nan                   # dropped
    cloud = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], axis=1)
    want = oracle.radius_outlier_removal(pts, radius, k)
    dp = gpu_ctx.radius_outlier_removal(cloud, radius, k)
    got = _fetch(dp)
    assert dp.shape[0] == int(want.sum())
    assert np.array_equal(got[:, :3], pts[want])          # survivors, input order preserved
    assert 0 < want.sum() < len(pts) or k == 0


def test_ror_generic_layout_and_empty(gpu_ctx, oracle):
    rng = np.random.default_rng(9)
    pts = _cloud(rng, 1500, 600)
    rec = np.zeros((len(pts), 8), np.float32)             # 32-byte XYZI-style records
    rec[:, 1], rec[:, 2], rec[:, 5] = pts[:, 0], pts[:, 1], pts[:, 2]
    want = oracle.radius_outlier_removal(pts, 0.2, 2)
    dp = gpu_ctx.radius_outlier_removal(rec, 0.2, 2, point_step=32, offsets=(4, 8, 20))
    assert np.array_equal(_fetch(dp)[:, :3], pts[want])
    dp = gpu_ctx.radius_outlier_removal(np.zeros((0, 4), np.float32))
    assert dp.shape[0] == 0


def test_ror_feeds_the_seed_stage(gpu_ctx, oracle):
    """globalMapCallback end to end: ROR on the device, its output straight into aos_seed_stage (no host copy)."""
    spec = synth.config("TINY", seed=5)
    spec.outlier_count = 40                # isolated in-window points: exactly what the filter removes
    pts = synth.make_orchard(spec)
    keep = oracle.radius_outlier_removal(pts, 0.2, 2)
    assert 0 < (~keep).sum()
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts[keep])
    dp = gpu_ctx.radius_outlier_removal(pts, 0.2, 2)
    assert dp.shape[0] == int(keep.sum())
    gpu_ctx.seed_stage(pl, dp)
    assert_seed_parity(gpu_ctx, r)
