"""Prototype of the chain-compressed BFS order (csrc/k_cluster.cu, bfs_chain_*): checks the event-driven walk against the
plain FIFO of seed_gen:1008-1049 on random blobs, thin curves, rings and oracle skeletons.  CPU only, pure Python.
    python scripts/dev/chain_bfs_proto.py
"""
import sys
import numpy as np

DX = (-1, -1, -1, 0, 0, 1, 1, 1)
DY = (-1, 0, 1, -1, 1, -1, 0, 1)


def components(img):
    from scipy import ndimage
    lab, n = ndimage.label(img, structure=np.ones((3, 3)))
    return lab, n


def plain_bfs(cells, root):
    """cells: dict (x, y) -> id; returns ids in FIFO order."""
    seen = {root}
    q = [root]
    h = 0
    while h < len(q):
        x, y = q[h]
        h += 1
        for k in range(8):
            p = (x + DX[k], y + DY[k])
            if p in cells and p not in seen:
                seen.add(p)
                q.append(p)
    return q


def chain_bfs(cells, root, stats=None):
    ids = {p: i for i, p in enumerate(cells)}
    pts = list(cells)
    n = len(pts)
    nbr = [[ids.get((x + DX[k], y + DY[k]), -1) for k in range(8)] for (x, y) in pts]
    r = ids[root]
    chain = [False] * n
    link = [None] * n
    for i in range(n):
        ks = [k for k in range(8) if nbr[i][k] >= 0]
        if len(ks) == 2 and i != r:
            a, b = ks
            if max(abs(DX[a] - DX[b]), abs(DY[a] - DY[b])) > 1:
                chain[i] = True
                link[i] = (nbr[i][a], nbr[i][b])
    # pointer jumping
    st = {}
    for i in range(n):
        if not chain[i]:
            continue
        for s in (0, 1):
            t = link[i][s]
            if chain[t]:
                ts = 1 if link[t][0] == i else 0
                st[(i, s)] = (t, ts, 1)
            else:
                st[(i, s)] = (i, s, 0)
    rounds = max(1, int(np.ceil(np.log2(max(n, 2)))) + 1)
    for _ in range(rounds):
        new = {}
        for (i, s), (t, ts, d) in st.items():
            if (t, ts) == (i, s):
                new[(i, s)] = (t, ts, d)
                continue
            t2, ts2, d2 = st[(t, ts)]
            new[(i, s)] = (t2, ts2, d + d2)
        st = new
    rb = [-1] * n
    pos = [0] * n
    L = [0] * n
    run_cells = []
    lohi = {}
    head_base = {}
    for i in range(n):
        if not chain[i]:
            continue
        e0, _, d0 = st[(i, 0)]
        e1, _, d1 = st[(i, 1)]
        L[i] = d0 + d1 + 1
        if e0 == e1:
            assert L[i] == 1, "pure ring of chain cells"
            head, pos[i] = i, 0
        elif e0 < e1:
            head, pos[i] = e0, d0
        else:
            head, pos[i] = e1, d1
        if head == i:
            assert pos[i] == 0
            head_base[i] = len(run_cells)
            run_cells.extend([-1] * L[i])
            lohi[head_base[i]] = [0, L[i]]
        rb[i] = head
    for i in range(n):
        if chain[i]:
            rb[i] = head_base[rb[i]]
            run_cells[rb[i] + pos[i]] = i
    assert all(c >= 0 for c in run_cells)
    vis = [False] * n

    def visited(j):
        if chain[j]:
            lo, hi = lohi[rb[j]]
            return not (lo <= pos[j] < hi)
        return vis[j]

    out = [r]
    vis[r] = True
    items = [(r, 0)]
    n_explicit = n_macro = 0
    while items:
        # macro step?
        delta = None
        for (u, d) in items:
            if not chain[u]:
                delta = 0
                break
            lo, hi = lohi[rb[u]]
            gap = hi - lo
            both = lo > 0 and hi < L[u]
            if L[u] == 1:
                f = 0
            elif both:
                f = gap // 2
            else:
                f = gap
            delta = f if delta is None else min(delta, f)
        if delta:
            n_macro += 1
            for t in range(1, delta + 1):
                for (u, d) in items:
                    out.append(run_cells[rb[u] + pos[u] + d * t])
            new_items = []
            for (u, d) in items:
                v = run_cells[rb[u] + pos[u] + d * delta]
                if d > 0:
                    lohi[rb[u]][0] = pos[v] + 1
                else:
                    lohi[rb[u]][1] = pos[v]
                new_items.append((v, d))
            items = new_items
            continue
        n_explicit += 1
        new_items = []
        for (u, d) in items:
            for k in range(8):
                j = nbr[u][k]
                if j < 0 or visited(j):
                    continue
                dj = 0
                if chain[j]:
                    if chain[u] and rb[u] == rb[j]:
                        dj = 1 if pos[u] == pos[j] - 1 else -1
                    else:
                        dj = 1 if pos[j] == 0 else -1
                    if dj > 0:
                        assert lohi[rb[j]][0] == pos[j]
                        lohi[rb[j]][0] = pos[j] + 1
                    else:
                        assert lohi[rb[j]][1] == pos[j] + 1
                        lohi[rb[j]][1] = pos[j]
                else:
                    vis[j] = True
                out.append(j)
                new_items.append((j, dj))
        items = new_items
    if stats is not None:
        stats.append((n, n_explicit, n_macro))
    return [pts[i] for i in out]


def check_image(img, name, stats=None):
    lab, ncomp = components(img)
    bad = 0
    for c in range(1, ncomp + 1):
        ys, xs = np.nonzero(lab == c)
        order = np.lexsort((xs, ys))
        cells = {(int(xs[i]), int(ys[i])): 1 for i in order}
        root = (int(xs[order[0]]), int(ys[order[0]]))
        a = plain_bfs(cells, root)
        b = chain_bfs(cells, root, stats)
        if a != b:
            bad += 1
            first = next(i for i in range(min(len(a), len(b))) if a[i] != b[i]) if len(a) == len(b) else -1
            print(f"{name}: component {c} ({len(cells)} cells) differs at {first}")
    return bad


def main():
    rng = np.random.default_rng(1)
    bad = 0
    for trial in range(300):
        h, w = rng.integers(3, 40, 2)
        img = rng.random((h, w)) < rng.choice([0.15, 0.3, 0.45, 0.6])
        bad += check_image(img, f"random{trial}")
    # thin curves: random walks, rings, crossings
    for trial in range(200):
        img = np.zeros((60, 120), bool)
        for _ in range(rng.integers(1, 5)):
            x, y = rng.integers(5, 115), rng.integers(5, 55)
            dxs = rng.choice([-1, 0, 1], p=[0.1, 0.1, 0.8]) if rng.random() < 0.5 else 1
            for _ in range(rng.integers(10, 150)):
                img[y % 60, x % 120] = True
                x += rng.choice([0, 1, 1, 1])
                y += rng.choice([-1, 0, 0, 0, 0, 1])
        if trial % 3 == 0:  # a ring
            cy, cx, rr = rng.integers(15, 45), rng.integers(20, 100), rng.integers(4, 14)
            t = np.linspace(0, 2 * np.pi, 400)
            img[(cy + rr * np.sin(t)).round().astype(int), (cx + rr * np.cos(t)).round().astype(int)] = True
        bad += check_image(img, f"curve{trial}")
    # pure shapes
    img = np.zeros((20, 20), bool)
    img[5, 3:15] = img[12, 3:15] = True
    img[5:13, 3] = img[5:13, 14] = True
    bad += check_image(img, "rectangle ring")
    img = np.zeros((30, 30), bool)
    for i in range(25):
        img[2 + i, 2 + i] = True
    bad += check_image(img, "diagonal")
    if len(sys.argv) > 1:
        sys.path[:0] = [".", "active-orchard-slam_b200"]
        from aos_gpu import synth
        from oracle import oracle
        spec = synth.config(sys.argv[1], seed=0)
        pts = synth.make_orchard(spec)
        oracle.set_fast(True, 8, skip_labels=True)
        p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius,
                              polygon=spec.polygon, exclusion=spec.exclusion)
        r = oracle.seed_stage(p, pts)
        stats = []
        bad += check_image(r["skel"] != 0, sys.argv[1], stats)
        stats.sort()
        for s in stats[-6:]:
            print("cells %d explicit levels %d macro steps %d" % s)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
