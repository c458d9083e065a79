"""Flat-array, scan-based model of DistributeOctTree — the executable specification of the CUDA
octree kernel (refactored_orb_slam2_b200/csrc/octree.cu).

The reference (ORBextractor.cc:529-731) walks a std::list and pushes children to the front; the
kernel cannot.  This model re-expresses every round as position arithmetic on arrays held in LIST
ORDER, which is what the kernel does with block-wide scans:

  full round    L' = reverse(children in creation order) ++ (single-key nodes in old order)
  careful round process the multi-key children of the previous round by (count desc, position asc)
                — position asc == creation sequence desc, the oracle's tie rule — until |L| >= N;
                L' = reverse(children of the processed nodes) ++ (unprocessed nodes in old order)

tests/test_octree_model.py checks it against the oracle's literal std::list restatement.
"""
import numpy as np

F = np.float32


def _round_half_away(v):
    return int(np.floor(v + 0.5)) if v >= 0 else -int(np.floor(-v + 0.5))


def distribute_flat(cands, min_x, max_x, min_y, max_y, n_want):
    """cands: int array [n,3] of (x, y, response) in octree coordinates, reference order.
    Returns the selected rows in final list order."""
    cands = np.asarray(cands, np.int64).reshape(-1, 3)
    n = len(cands)
    if n == 0:
        return cands
    cx, cy, cr = cands[:, 0], cands[:, 1], cands[:, 2]
    n_ini = _round_half_away(float(F(max_x - min_x) / F(max_y - min_y)))
    hx = F(max_x - min_x) / F(n_ini)
    root = (cx.astype(F) / hx).astype(np.int64)
    # roots in order, empty ones erased
    rx0 = np.array([int(hx * F(i)) for i in range(n_ini)], np.int64)
    rx1 = np.array([int(hx * F(i + 1)) for i in range(n_ini)], np.int64)
    cnt_root = np.bincount(root, minlength=n_ini)
    alive = cnt_root > 0
    pos_of_root = np.cumsum(alive) - 1
    X0, X1 = rx0[alive], rx1[alive]
    Y0 = np.zeros(alive.sum(), np.int64)
    Y1 = np.full(alive.sum(), max_y - min_y, np.int64)
    CNT = cnt_root[alive]
    node = pos_of_root[root]

    def split_info(sel):
        """For the nodes listed in `sel` (positions): midpoints, per-child counts."""
        mx = X0[sel] + (X1[sel] - X0[sel] + 1) // 2   # ceil((x1-x0)/2.f) for non-negative ints
        my = Y0[sel] + (Y1[sel] - Y0[sel] + 1) // 2
        slot = np.full(len(CNT), -1, np.int64)
        slot[sel] = np.arange(len(sel))
        cs = slot[node]
        m = cs >= 0
        quad = np.zeros(n, np.int64)
        quad[m] = np.where(cx[m] < mx[cs[m]], np.where(cy[m] < my[cs[m]], 0, 2), np.where(cy[m] < my[cs[m]], 1, 3))
        cc = np.zeros((len(sel), 4), np.int64)
        np.add.at(cc, (cs[m], quad[m]), 1)
        return mx, my, slot, quad, cc

    def apply_split(sel, nproc, order_is_sel):
        """Split sel[:nproc] (already in processing order); everything else keeps old relative order."""
        nonlocal X0, X1, Y0, Y1, CNT, node
        proc = sel[:nproc]
        mx, my, slot, quad, cc = split_info(proc)
        ne = cc > 0
        nch = ne.sum(axis=1)
        base = np.cumsum(nch) - nch
        T = int(nch.sum())
        within = np.cumsum(ne, axis=1) - ne          # non-empty children before q
        child_pos = T - 1 - (base[:, None] + within)  # valid where ne
        keep = np.ones(len(CNT), bool)
        keep[proc] = False
        keep_rank = np.cumsum(keep) - 1
        new_size = T + int(keep.sum())
        nX0, nX1, nY0, nY1, nC = (np.zeros(new_size, np.int64) for _ in range(5))
        kp = T + keep_rank[keep]
        nX0[kp], nX1[kp], nY0[kp], nY1[kp], nC[kp] = X0[keep], X1[keep], Y0[keep], Y1[keep], CNT[keep]
        px0, px1, py0, py1 = X0[proc], X1[proc], Y0[proc], Y1[proc]
        rects = [(px0, mx, py0, my), (mx, px1, py0, my), (px0, mx, my, py1), (mx, px1, my, py1)]
        for q in range(4):
            s = ne[:, q]
            p = child_pos[s, q]
            nX0[p], nX1[p], nY0[p], nY1[p] = rects[q][0][s], rects[q][1][s], rects[q][2][s], rects[q][3][s]
            nC[p] = cc[s, q]
        cs = slot[node]
        m = cs >= 0
        new_node = np.empty(n, np.int64)
        new_node[m] = child_pos[cs[m], quad[m]]
        new_node[~m] = T + keep_rank[node[~m]]
        X0, X1, Y0, Y1, CNT, node = nX0, nX1, nY0, nY1, nC, new_node
        return T

    finish = False
    while not finish:
        prev_size = len(CNT)
        multi = np.nonzero(CNT > 1)[0]
        T = apply_split(multi, len(multi), True)
        size = len(CNT)
        n_to_expand = int((CNT[:T] > 1).sum())
        if size >= n_want or size == prev_size:
            finish = True
        elif size + 3 * n_to_expand > n_want:
            while not finish:
                prev_size = len(CNT)
                elig = np.nonzero(CNT[:T] > 1)[0]
                # (count desc, position asc)
                order = elig[np.lexsort((elig, -CNT[elig]))]
                if len(order) == 0:
                    break
                _, _, _, _, cc = split_info(order)
                gain = (cc > 0).sum(axis=1) - 1
                run = prev_size + np.cumsum(gain)
                hit = np.nonzero(run >= n_want)[0]
                nproc = int(hit[0]) + 1 if len(hit) else len(order)
                T = apply_split(order, nproc, False)
                size = len(CNT)
                if size >= n_want or size == prev_size:
                    finish = True
    # first key (in reference order) with maximal response per node
    out = np.zeros((len(CNT), 3), np.int64)
    best = np.full(len(CNT), -1, np.int64)
    for k in range(n):
        b = best[node[k]]
        if b < 0 or cr[k] > cr[b]:
            best[node[k]] = k
    out[:] = cands[best]
    return out
