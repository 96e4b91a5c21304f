"""Seeded synthetic orchards: the reference ships no sample data.

Rows run parallel to x at `row_pitch`, trees every `tree_spacing` m along a row; each tree is a
vertical cylinder of radius `tree_radius` with z ~ U(-1, 3) so the z window [-0.4, 0.5] keeps
about 22 % of its points.  Ground clutter sits outside the z window, a few in-window outliers make
small blobs that the cluster-length filter has to reject.  The exploration polygon is the extent
inset by 2.5 m, so the +-2.5 m active bounds (src/aos_seed_gen_node.cpp:874-890) reproduce the
extent and the BASELINE.json grid sizes (100x60 m -> 2000x1200 @0.05 m, 1 km^2 -> 20000^2).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# aos_seed_gen_node's 11 hard-coded exclusion discs (src/aos_seed_gen_node.cpp:487-499): field
# constants of the authors' orchard; passed as a parameter array here.
REFERENCE_EXCLUSION_DISCS = np.array(
    [[0.646417, 3.83918, 1.0], [2.0405, 3.62485, 1.0], [65.3711, 2.09755, 1.0], [66.9094, 2.07515, 1.0],
     [-1.61309, 5.69933, 1.0], [-1.97349, 4.77329, 1.0], [-2.11365, 3.74464, 1.0], [-2.26381, 2.70848, 1.0],
     [-2.66426, 1.72738, 1.0], [68.0229, 2.31687, 1.0], [65.4647, 2.18653, 1.0]], dtype=np.float32)

# the default exploration polygon (src/aos_seed_gen_node.cpp:196-199)
REFERENCE_POLYGON = np.array(
    [[-1.972916603088379, 7.9420671463012695], [-2.0726776123046875, 0.022441387176513672],
     [70.22465515136719, 2.102720260620117], [69.48777770996094, 9.786612510681152]], dtype=np.float64)


@dataclass
class OrchardSpec:
    extent_x: float = 100.0
    extent_y: float = 60.0
    origin_x: float = 0.0          # world coordinate of the extent's lower-left corner
    origin_y: float = 0.0
    row_pitch: float = 6.0
    tree_spacing: float = 2.5
    tree_radius: float = 0.5
    n_points: int = 2_000_000
    gap_prob: float = 0.02         # probability that a tree is missing (breaks rows into clusters)
    jitter: float = 0.10           # trunk position jitter (m)
    clutter_frac: float = 0.08     # ground / canopy clutter outside the z window
    outlier_count: int = 12        # isolated in-window points
    seed: int = 0
    rotation_deg: float = 0.0      # rows rotated about the centre of the extent (0 = parallel to x)
    grid_resolution: float = 0.05
    inflation_radius: float = 0.8
    exclusion: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))

    @property
    def polygon(self) -> np.ndarray:
        x0, y0 = self.origin_x + 2.5, self.origin_y + 2.5
        x1, y1 = self.origin_x + self.extent_x - 2.5, self.origin_y + self.extent_y - 2.5
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.float64)


def config(name: str, seed: int = 0, n_points: int | None = None) -> OrchardSpec:
    """BASELINE.json configs: C1 1000x600 @0.1 m, C2 2000x1200 @0.05 m, C3 20000^2 @0.05 m,
    C4 40000^2 @0.025 m, plus small ones for unit tests."""
    name = name.upper()
    if name == "C1":
        s = OrchardSpec(grid_resolution=0.1, seed=seed)
    elif name == "C2":
        s = OrchardSpec(seed=seed)
    elif name == "C3":
        s = OrchardSpec(extent_x=1000.0, extent_y=1000.0, row_pitch=4.0, n_points=200_000_000, seed=seed)
    elif name == "C4":
        s = OrchardSpec(extent_x=1000.0, extent_y=1000.0, row_pitch=4.0, n_points=200_000_000,
                        grid_resolution=0.025, seed=seed)
    elif name == "TINY":   # 400 x 240 cells
        s = OrchardSpec(extent_x=20.0, extent_y=12.0, row_pitch=4.0, n_points=40_000, outlier_count=3, seed=seed)
    elif name == "SMALL":  # 800 x 480 cells
        s = OrchardSpec(extent_x=40.0, extent_y=24.0, n_points=200_000, outlier_count=6, seed=seed)
    else:
        raise ValueError(name)
    if n_points is not None:
        s.n_points = n_points
    return s


def tree_centres(spec: OrchardSpec, rng: np.random.Generator) -> np.ndarray:
    n_rows = int(np.floor(spec.extent_y / spec.row_pitch))
    n_trees = int(np.floor(spec.extent_x / spec.tree_spacing))
    ys = spec.origin_y + spec.row_pitch * (0.5 + np.arange(n_rows))
    xs = spec.origin_x + spec.tree_spacing * (0.5 + np.arange(n_trees))
    cx, cy = np.meshgrid(xs, ys)
    c = np.stack([cx.ravel(), cy.ravel()], axis=1)
    c += rng.uniform(-spec.jitter, spec.jitter, size=c.shape)
    keep = rng.random(len(c)) >= spec.gap_prob
    c = c[keep]
    if spec.rotation_deg != 0.0:   # rotate the layout, keep the trees that stay well inside the extent
        a = np.deg2rad(spec.rotation_deg)
        mid = np.array([spec.origin_x + 0.5 * spec.extent_x, spec.origin_y + 0.5 * spec.extent_y])
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        c = (c - mid) @ rot.T + mid
        m = 3.0
        inside = ((c[:, 0] > spec.origin_x + m) & (c[:, 0] < spec.origin_x + spec.extent_x - m) &
                  (c[:, 1] > spec.origin_y + m) & (c[:, 1] < spec.origin_y + spec.extent_y - m))
        c = c[inside]
    return c


def make_orchard(spec: OrchardSpec) -> np.ndarray:
    """float32 [N, 4] points (x, y, z, pad) -- PointXYZ's 16-byte layout."""
    rng = np.random.default_rng(spec.seed)
    centres = tree_centres(spec, rng)
    n_clutter = int(spec.n_points * spec.clutter_frac)
    n_tree_pts = spec.n_points - n_clutter - spec.outlier_count
    tree_of = rng.integers(0, len(centres), size=n_tree_pts)
    r = spec.tree_radius * np.sqrt(rng.random(n_tree_pts))
    th = rng.uniform(0.0, 2.0 * np.pi, n_tree_pts)
    pts = np.empty((spec.n_points, 4), np.float32)
    pts[:n_tree_pts, 0] = centres[tree_of, 0] + r * np.cos(th)
    pts[:n_tree_pts, 1] = centres[tree_of, 1] + r * np.sin(th)
    pts[:n_tree_pts, 2] = rng.uniform(-1.0, 3.0, n_tree_pts)
    a, b = n_tree_pts, n_tree_pts + n_clutter
    pts[a:b, 0] = rng.uniform(spec.origin_x - 1.0, spec.origin_x + spec.extent_x + 1.0, n_clutter)
    pts[a:b, 1] = rng.uniform(spec.origin_y - 1.0, spec.origin_y + spec.extent_y + 1.0, n_clutter)
    zc = rng.uniform(0.0, 1.0, n_clutter)
    pts[a:b, 2] = np.where(zc < 0.5, -1.5 + 1.0 * zc * 2.0, 0.6 + 2.4 * (zc - 0.5) * 2.0)  # (-1.5,-0.5) u (0.6,3.0)
    pts[b:, 0] = rng.uniform(spec.origin_x, spec.origin_x + spec.extent_x, spec.outlier_count)
    pts[b:, 1] = rng.uniform(spec.origin_y, spec.origin_y + spec.extent_y, spec.outlier_count)
    pts[b:, 2] = rng.uniform(-0.3, 0.4, spec.outlier_count)
    pts[:, 3] = 1.0
    rng.shuffle(pts, axis=0)
    return pts


def make_orchard_torch(spec: OrchardSpec, device, chunk: int = 1 << 24, y_range=None):
    """Same recipe generated on `device` with torch (for the 200 M-point C3/C4 clouds).
    Not bit-identical to make_orchard(); full-size parity goes through size-independent properties.
    y_range=(lo, hi): only the share of the cloud whose trees / clutter fall into that strip (row-band sharding:
    every GPU generates the points of its own rows; the tree layout is the same global one)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(spec.seed if y_range is None else spec.seed * 1000003 + int(y_range[0] * 16))
    rng = np.random.default_rng(spec.seed)
    all_centres = tree_centres(spec, rng)
    n = spec.n_points
    n_clutter = int(n * spec.clutter_frac)
    n_tree_pts = n - n_clutter
    cy0, cy1 = spec.origin_y - 1.0, spec.origin_y + spec.extent_y + 1.0
    if y_range is not None:
        lo, hi = y_range
        sel = (all_centres[:, 1] >= lo - spec.tree_radius) & (all_centres[:, 1] <= hi + spec.tree_radius)
        n_tree_pts = int(round(n_tree_pts * sel.sum() / max(len(all_centres), 1)))
        all_centres = all_centres[sel]
        cy0, cy1 = max(cy0, lo), min(cy1, hi)
        n_clutter = int(round(n_clutter * max(cy1 - cy0, 0.0) / (spec.extent_y + 2.0)))
        n = n_tree_pts + n_clutter
    if len(all_centres) == 0:
        n_tree_pts, n = 0, n_clutter
    centres = torch.from_numpy(all_centres).to(device=device, dtype=torch.float32)
    out = torch.empty((n, 4), dtype=torch.float32, device=device)
    for s in range(0, n_tree_pts, chunk):
        m = min(chunk, n_tree_pts - s)
        idx = torch.randint(0, centres.shape[0], (m,), generator=g, device=device)
        r = spec.tree_radius * torch.sqrt(torch.rand(m, generator=g, device=device))
        th = 2.0 * np.pi * torch.rand(m, generator=g, device=device)
        out[s:s + m, 0] = centres[idx, 0] + r * torch.cos(th)
        out[s:s + m, 1] = centres[idx, 1] + r * torch.sin(th)
        out[s:s + m, 2] = -1.0 + 4.0 * torch.rand(m, generator=g, device=device)
    for s in range(n_tree_pts, n, chunk):
        m = min(chunk, n - s)
        out[s:s + m, 0] = spec.origin_x - 1.0 + (spec.extent_x + 2.0) * torch.rand(m, generator=g, device=device)
        out[s:s + m, 1] = cy0 + (cy1 - cy0) * torch.rand(m, generator=g, device=device)
        zc = torch.rand(m, generator=g, device=device)
        out[s:s + m, 2] = torch.where(zc < 0.5, -1.5 + 2.0 * zc, 0.6 + 4.8 * (zc - 0.5))
    out[:, 3] = 1.0
    return out


N_STRIPS = 64


def make_orchard_strips_torch(spec: OrchardSpec, device, y_range=None, n_strips: int = N_STRIPS, chunk: int = 1 << 24):
    """The orchard cloud generated in `n_strips` fixed horizontal strips, each from its OWN generator seed, so that the
    union of the strips is one global cloud whatever subset a process generates: row-band sharding on N GPUs (every
    rank generates the strips that touch its rows, y_range=(lo, hi)) then works on exactly the points a single GPU
    gets with y_range=None, and the results can be compared digest for digest at N = 1, 2, 4, 8.
    Strip s holds the trees whose centre y lies in it (all their points) and the clutter with y in it; the point count
    of a strip is fixed by the spec alone."""
    import torch

    rng = np.random.default_rng(spec.seed)
    centres_all = tree_centres(spec, rng)
    y0, y1 = spec.origin_y - 1.0, spec.origin_y + spec.extent_y + 1.0
    edges = np.linspace(y0, y1, n_strips + 1)
    strip_of = np.clip(np.searchsorted(edges, centres_all[:, 1], side="right") - 1, 0, n_strips - 1)
    n_clutter = int(spec.n_points * spec.clutter_frac)
    n_tree_pts = spec.n_points - n_clutter
    per_tree = n_tree_pts // max(len(centres_all), 1)
    clutter_per_strip = n_clutter // n_strips
    parts = []
    for s in range(n_strips):
        lo, hi = float(edges[s]), float(edges[s + 1])
        if y_range is not None and (hi + spec.tree_radius + 0.01 < y_range[0] or lo - spec.tree_radius - 0.01 > y_range[1]):
            continue
        g = torch.Generator(device=device)
        g.manual_seed(spec.seed * 1000003 + 7919 * s + 17)
        c = torch.from_numpy(centres_all[strip_of == s]).to(device=device, dtype=torch.float32)
        m = per_tree * c.shape[0]
        out = torch.empty((m + clutter_per_strip, 4), dtype=torch.float32, device=device)
        for a in range(0, m, chunk):
            k = min(chunk, m - a)
            idx = torch.randint(0, c.shape[0], (k,), generator=g, device=device)
            r = spec.tree_radius * torch.sqrt(torch.rand(k, generator=g, device=device))
            th = 2.0 * np.pi * torch.rand(k, generator=g, device=device)
            out[a:a + k, 0] = c[idx, 0] + r * torch.cos(th)
            out[a:a + k, 1] = c[idx, 1] + r * torch.sin(th)
            out[a:a + k, 2] = -1.0 + 4.0 * torch.rand(k, generator=g, device=device)
        k = clutter_per_strip
        out[m:, 0] = spec.origin_x - 1.0 + (spec.extent_x + 2.0) * torch.rand(k, generator=g, device=device)
        out[m:, 1] = lo + (hi - lo) * torch.rand(k, generator=g, device=device)
        zc = torch.rand(k, generator=g, device=device)
        out[m:, 2] = torch.where(zc < 0.5, -1.5 + 2.0 * zc, 0.6 + 4.8 * (zc - 0.5))
        out[:, 3] = 1.0
        parts.append(out)
    if not parts:
        return torch.empty((0, 4), dtype=torch.float32, device=device)
    return torch.cat(parts, dim=0)


def make_orchard_strips(spec: OrchardSpec, n_strips: int = N_STRIPS, threads: int = 0) -> np.ndarray:
    """Full-size clouds on the CPU in seconds, bit-identical on every machine: the orchard recipe of make_orchard()
    generated strip by strip (one numpy Generator per strip, strips in parallel threads) using only IEEE-exact
    operations -- points are drawn in the square around a trunk and kept when x^2 + y^2 <= r^2 (no sin/cos, whose
    vectorised implementations differ between CPUs), so digests of results on these clouds can be committed as golden
    values (tests/golden/fullsize_digests.json).  Points are ordered strip by strip, tree-interleaved inside a strip."""
    import concurrent.futures as cf
    import os

    rng0 = np.random.default_rng(spec.seed)
    centres_all = tree_centres(spec, rng0).astype(np.float32)
    y0, y1 = spec.origin_y - 1.0, spec.origin_y + spec.extent_y + 1.0
    edges = np.linspace(y0, y1, n_strips + 1)
    strip_of = np.clip(np.searchsorted(edges, centres_all[:, 1], side="right") - 1, 0, n_strips - 1)
    n_clutter = int(spec.n_points * spec.clutter_frac)
    n_tree_pts = spec.n_points - n_clutter
    per_tree = n_tree_pts // max(len(centres_all), 1)
    clutter_per_strip = n_clutter // n_strips
    rad = np.float32(spec.tree_radius)

    def one(s):
        rng = np.random.default_rng([spec.seed, 7919, s])
        c = centres_all[strip_of == s]
        m = per_tree * len(c)
        out = np.empty((m + clutter_per_strip, 4), np.float32)
        got = 0
        if m:
            cand = int(m * 1.30) + 4096
            u = rng.random((cand, 2), dtype=np.float32) * np.float32(2.0) - np.float32(1.0)
            keep = np.nonzero(u[:, 0] * u[:, 0] + u[:, 1] * u[:, 1] <= np.float32(1.0))[0][:m]
            got = len(keep)
            idx = rng.integers(0, len(c), size=got)
            out[:got, 0] = c[idx, 0] + rad * u[keep, 0]
            out[:got, 1] = c[idx, 1] + rad * u[keep, 1]
            out[:got, 2] = rng.random(got, dtype=np.float32) * np.float32(4.0) - np.float32(1.0)
        k = clutter_per_strip
        lo, hi = np.float32(edges[s]), np.float32(edges[s + 1])
        out[got:got + k, 0] = np.float32(spec.origin_x - 1.0) + np.float32(spec.extent_x + 2.0) * rng.random(k, dtype=np.float32)
        out[got:got + k, 1] = lo + (hi - lo) * rng.random(k, dtype=np.float32)
        zc = rng.random(k, dtype=np.float32)
        out[got:got + k, 2] = np.where(zc < np.float32(0.5), np.float32(-1.5) + np.float32(2.0) * zc,
                                       np.float32(0.6) + np.float32(4.8) * (zc - np.float32(0.5)))
        out[:, 3] = 1.0
        return out[:got + k]

    with cf.ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 1)) as pool:
        parts = list(pool.map(one, range(n_strips)))
    total = sum(len(p) for p in parts)
    pts = np.empty((total, 4), np.float32)
    a = 0
    for p in parts:
        pts[a:a + len(p)] = p
        a += len(p)
    return pts
