"""ORACLE (test infrastructure): an independent numpy/float32 restatement of the
reference's binned-SAH builder, Util/BVH.fs:128-247, used to
cross-check the C++ host-side builder the product ships (bn_host_bvh_build).
Pure Python recursion over numpy slices; fine for ~10^5 items.

PARITY UNPINNED by the reference (it has no tests); pinned here by agreement of two
independently written implementations plus the structural invariants in
tests/test_bvh_build.py.
"""
from __future__ import annotations

import numpy as np

F = np.float32
SAH_BINS, MAX_LEAF, MAX_DEPTH = 12, 4, 64  # BVHBuildConfig, Util/BVH.fs:99-107


def _surface_area(lo, hi):
    d = hi - lo
    return F(2) * ((d[0] * d[1] + d[1] * d[2]) + d[2] * d[0])  # Util/BVH.fs:17-19


def build(boxes: np.ndarray):
    """boxes: [n, 6] float32 (min xyz, max xyz).  Returns (nodes, perm) where nodes is a
    list of dicts in preorder {lo, hi, leaf, first/count | right, axis} and perm[i] is the
    original index of the item at slot i."""
    boxes = np.ascontiguousarray(boxes, dtype=F)
    n = boxes.shape[0]
    lo_all, hi_all = boxes[:, :3].copy(), boxes[:, 3:].copy()
    order = np.arange(n)
    nodes = []

    def rec(first, last, depth):
        idx = order[first:last].copy()          # the saved copy, :130
        lo, hi = lo_all[idx], hi_all[idx]
        count = last - first
        blo, bhi = lo.min(axis=0), hi.max(axis=0)
        me = len(nodes)
        if count <= MAX_LEAF or depth >= MAX_DEPTH:
            nodes.append({"lo": blo, "hi": bhi, "leaf": True, "first": first, "count": count})
            return me
        cen = F(0.5) * (lo + hi)                # Centroid, :14
        clo, chi = cen.min(axis=0), cen.max(axis=0)
        d = chi - clo
        axis = 0 if (d[0] >= d[1] and d[0] >= d[2]) else (1 if d[1] >= d[2] else 2)  # :24-27
        extent = d[axis]
        if extent == 0:
            mid = first + count // 2            # :150-156
        else:
            with np.errstate(invalid="ignore", over="ignore"):
                b = np.minimum(((F(SAH_BINS) * (cen[:, axis] - clo[axis])) / extent).astype(np.int32), SAH_BINS - 1)  # :163-169
                inf = F(np.inf)
                bin_lo = np.full((SAH_BINS, 3), inf, dtype=F)
                bin_hi = np.full((SAH_BINS, 3), -inf, dtype=F)
                bin_n = np.zeros(SAH_BINS, dtype=np.int64)
                for k in range(SAH_BINS):
                    m = b == k
                    if m.any():
                        bin_lo[k], bin_hi[k], bin_n[k] = lo[m].min(axis=0), hi[m].max(axis=0), m.sum()
                nc = SAH_BINS - 1
                costs = np.zeros(nc, dtype=F)
                l_lo, l_hi, l_n = np.full(3, inf, F), np.full(3, -inf, F), 0
                r_lo, r_hi, r_n = np.full(3, inf, F), np.full(3, -inf, F), 0
                for i in range(nc):             # :187-193
                    l_lo, l_hi, l_n = np.minimum(l_lo, bin_lo[i]), np.maximum(l_hi, bin_hi[i]), l_n + bin_n[i]
                    costs[i] = costs[i] + F(l_n) * _surface_area(l_lo, l_hi)
                    j = nc - i
                    r_lo, r_hi, r_n = np.minimum(r_lo, bin_lo[j]), np.maximum(r_hi, bin_hi[j]), r_n + bin_n[j]
                    costs[nc - 1 - i] = costs[nc - 1 - i] + F(r_n) * _surface_area(r_lo, r_hi)
                base = F(count) * _surface_area(clo, chi)
                best, min_cost = 0, inf
                for i in range(nc):             # :198-203, strict <
                    c = costs[i] + base
                    if c < min_cost:
                        min_cost, best = c, i
            left_mask = b <= best
            left_items = idx[left_mask]
            right_items = idx[~left_mask][::-1]   # right side is filled from the end, :212-214
            order[first:first + len(left_items)] = left_items
            order[first + len(left_items):last] = right_items
            mid = first + len(left_items)
        nodes.append(None)
        rec(first, mid, depth + 1)
        right = rec(mid, last, depth + 1)
        nodes[me] = {"lo": blo, "hi": bhi, "leaf": False, "right": right, "axis": axis}
        return me

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    rec(0, n, 0)
    return nodes, order.copy()
