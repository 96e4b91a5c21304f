"""Tolerance comparison of two GvdGraph results of the same map (used for the opt-in device Voronoi, whose graph equals the
reference's only up to Subdiv2D's history-dependent choices; the default path is compared bit for bit, not with this)."""
from __future__ import annotations

import numpy as np


def compare_graphs(g_ref, g_dev, tol=1e-4):
    """Fractions of agreement between two GvdGraph dicts (replay = reference-exact, device = opt-in)."""
    from scipy.spatial import cKDTree
    a, b = g_ref["nodes"], g_dev["nodes"]
    ta, tb = cKDTree(a), cKDTree(b)
    d_ab, i_ab = tb.query(a)            # for every reference node: nearest device node
    d_ba, _ = ta.query(b)
    m_ab = d_ab <= tol
    out = {"nodes_ref": len(a), "nodes_dev": len(b), "ref_nodes_matched": float(m_ab.mean()), "dev_nodes_matched": float((d_ba <= tol).mean()),
           "ref_nodes_within_1mm": float((d_ab <= 1e-3).mean()),
           "bitwise_equal_of_matched": float((a[m_ab] == b[i_ab[m_ab]]).all(axis=1).mean()),
           "max_err_matched_m": float(d_ab[m_ab].max()) if m_ab.any() else None,
           "worst_unmatched_m": float(d_ab.max())}
    # edges of the reference graph whose two end nodes are matched, looked up in the device graph
    eb = set(map(tuple, np.sort(g_dev["edges"], axis=1).tolist()))
    ea = g_ref["edges"]
    both = m_ab[ea[:, 0]] & m_ab[ea[:, 1]]
    mapped = np.sort(np.stack([i_ab[ea[both, 0]], i_ab[ea[both, 1]]], axis=1), axis=1)
    hit = sum(1 for e in map(tuple, mapped.tolist()) if e in eb)
    out.update(edges_ref=len(ea), edges_dev=len(g_dev["edges"]), ref_edges_with_matched_ends=float(both.mean()),
               ref_edges_found=hit / max(len(ea), 1))
    # labelled nodes (TL/TR/BL/BR): same points with the same label masks
    la = {tuple(np.round(a[i], 3)): int(g_ref["node_labels"][i]) for i in np.nonzero(g_ref["node_labels"])[0]}
    lb = {tuple(np.round(b[i], 3)): int(g_dev["node_labels"][i]) for i in np.nonzero(g_dev["node_labels"])[0]}
    same = sum(1 for k, v in la.items() if lb.get(k) == v)
    out.update(labelled_ref=len(la), labelled_same=same / max(len(la), 1))
    return out
