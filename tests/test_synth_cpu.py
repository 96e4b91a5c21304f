"""Synthetic-cloud generators (no GPU): the committed full-size digests and the N-invariance of the band bench rest on them."""
import hashlib

import numpy as np

from aos_gpu import synth

# sha256 of make_orchard_strips(config("C3", seed=3, n_points=2_000_000)): IEEE-exact operations only, so every machine
# regenerates the golden clouds of tests/golden/fullsize_digests.json bit for bit
GOLDEN_CLOUD_2M = "955bfe5e0b74fd0f1be394ff5d08588f2090920ec0412234ed2495126d088d21"


def test_strip_cloud_is_reproducible_and_thread_independent():
    spec = synth.config("C3", seed=3, n_points=2_000_000)
    a = synth.make_orchard_strips(spec, threads=1)
    b = synth.make_orchard_strips(spec, threads=5)
    assert np.array_equal(a, b)
    assert hashlib.sha256(a.tobytes()).hexdigest() == GOLDEN_CLOUD_2M
    assert np.isfinite(a).all() and abs(len(a) - spec.n_points) < 0.05 * spec.n_points   # whole points per tree
    keep = (a[:, 2] >= -0.4) & (a[:, 2] <= 0.5)
    assert 0.15 < keep.mean() < 0.3          # the z window keeps about a fifth, as the generator spec says


def test_band_strips_are_subsets_of_one_global_cloud():
    """make_orchard_strips_torch(y_range=...) must return whole strips of the SAME global cloud whatever the split, which
    is what lets bench.py compare the band run's digest with the single-GPU digest at every N."""
    import torch
    spec = synth.OrchardSpec(extent_x=60.0, extent_y=48.0, row_pitch=4.0, n_points=120_000, seed=2)
    full = synth.make_orchard_strips_torch(spec, torch.device("cpu"), n_strips=8)
    full_rows = {r.tobytes() for r in full.numpy()}
    assert len(full_rows) > 0.99 * len(full)
    covered = set()
    for lo, hi in ((-1.0, 15.0), (15.0, 33.0), (33.0, 49.0)):
        part = synth.make_orchard_strips_torch(spec, torch.device("cpu"), y_range=(lo, hi), n_strips=8)
        rows = {r.tobytes() for r in part.numpy()}
        assert rows <= full_rows                       # nothing a single GPU would not also get
        # every point of the full cloud that lies inside the band's rows is there
        f = full.numpy()
        inside = f[(f[:, 1] >= lo) & (f[:, 1] <= hi)]
        assert {r.tobytes() for r in inside} <= rows
        covered |= rows
    assert covered == full_rows
