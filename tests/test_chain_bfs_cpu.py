"""The BFS order the library needs for order-dependent clusters (seed_gen:1008-1059) is produced without walking cell by
cell (csrc/k_cluster.cu, bfs_chain_kernel): chain cells are ranked into runs and levels of chain walkers are fast-forwarded.
scripts/dev/chain_bfs_proto.py states that algorithm in plain Python; this test checks it against the literal FIFO on
random blobs, thin random walks, rings and crossings.  (The CUDA kernel itself is checked against the oracle by the -m gpu
tests: test_seed_gpu.py, test_fullsize_exact_gpu.py.)"""
import importlib.util
import os

import numpy as np

_P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "dev", "chain_bfs_proto.py")
_spec = importlib.util.spec_from_file_location("chain_bfs_proto", _P)
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


def test_random_blobs():
    rng = np.random.default_rng(7)
    for trial in range(80):
        h, w = rng.integers(3, 32, 2)
        img = rng.random((h, w)) < rng.choice([0.15, 0.3, 0.45, 0.6])
        assert proto.check_image(img, f"random{trial}") == 0


def test_thin_curves_rings_and_crossings():
    rng = np.random.default_rng(8)
    stats = []
    for trial in range(40):
        img = np.zeros((60, 120), bool)
        for _ in range(rng.integers(1, 5)):
            x, y = rng.integers(5, 115), rng.integers(5, 55)
            for _ in range(rng.integers(10, 150)):
                img[y % 60, x % 120] = True
                x += rng.choice([0, 1, 1, 1])
                y += rng.choice([-1, 0, 0, 0, 0, 1])
        if trial % 3 == 0:
            cy, cx, rr = rng.integers(15, 45), rng.integers(20, 100), rng.integers(4, 14)
            t = np.linspace(0, 2 * np.pi, 400)
            img[(cy + rr * np.sin(t)).round().astype(int), (cx + rr * np.cos(t)).round().astype(int)] = True
        assert proto.check_image(img, f"curve{trial}", stats) == 0
    # the point of the exercise: far fewer sequential steps than cells on thin shapes
    cells = sum(s[0] for s in stats)
    steps = sum(s[1] + s[2] for s in stats)
    assert steps < cells


def test_straight_shapes():
    img = np.zeros((20, 20), bool)
    img[5, 3:15] = img[12, 3:15] = True
    img[5:13, 3] = img[5:13, 14] = True
    assert proto.check_image(img, "rectangle ring") == 0
    img = np.zeros((30, 30), bool)
    for i in range(25):
        img[2 + i, 2 + i] = True
    assert proto.check_image(img, "diagonal") == 0
    img = np.zeros((3, 400), bool)
    img[1, :] = True
    st = []
    assert proto.check_image(img, "line", st) == 0
    assert st[0][1] + st[0][2] < 12   # a 400-cell line takes a handful of steps
