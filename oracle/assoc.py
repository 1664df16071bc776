"""TEST INFRASTRUCTURE ONLY (checker, never shipped / never imported by fastmot_b200).

Numpy/pure-Python restatement of the reference association primitives:
  rect helpers        fastmot/utils/rect.py:5-57, 100-109, 142-157
  cdist / iou_dist    fastmot/utils/distance.py:16-108
  fuse / gate         fastmot/utils/matching.py:100-116
  linear_assignment   fastmot/utils/matching.py:10-30, 57-70  (+ SciPy 1.18.1 rectangular_lsap, restated below
                      because SciPy is a third-party dependency of the reference: requirements.txt:2)
  greedy_match        fastmot/utils/matching.py:33-54, 73-97
"""
import numpy as np

INF_COST = 1e5
CHI_SQ_INV_95 = 9.4877


def round_half_even(x):
    return np.rint(np.asarray(x, np.float64))


def to_tlbr(tlwh):
    t = np.asarray(tlwh, np.float64)
    return np.rint(np.stack([t[..., 0], t[..., 1], t[..., 0] + t[..., 2] - 1., t[..., 1] + t[..., 3] - 1.], -1))


def area(tlbr):
    t = np.asarray(tlbr, np.float64)
    w, h = t[..., 2] - t[..., 0] + 1, t[..., 3] - t[..., 1] + 1
    return np.where((w <= 0) | (h <= 0), 0., w * h)


def _inter(a, b):
    iw = np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]) + 1
    ih = np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]) + 1
    return iw, ih


def ios(tlbr, rect):
    a = np.asarray(tlbr, np.float64).reshape(-1, 4)
    iw, ih = _inter(a, np.asarray(rect, np.float64).reshape(1, 4))
    iw, ih = iw[:, 0], ih[:, 0]
    with np.errstate(divide='ignore', invalid='ignore'):
        v = iw * ih / area(a)
    return np.where((iw <= 0) | (ih <= 0), 0., v)


def iou_dist(a, b):
    a = np.asarray(a, np.float64).reshape(-1, 4)
    b = np.asarray(b, np.float64).reshape(-1, 4)
    iw, ih = _inter(a, b)
    inter = iw * ih
    union = area(a)[:, None] + area(b)[None, :] - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        d = 1. - inter / union
    return np.where((iw > 0) & (ih > 0), d, 1.)


def find_occluded(tlbr, thresh):
    a = np.asarray(tlbr, np.float64).reshape(-1, 4)
    n = len(a)
    if n == 0:
        return np.zeros(0, bool)
    iw, ih = _inter(a, a)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = iw * ih / area(a)[:, None]
    hit = (iw > 0) & (ih > 0) & (r >= thresh)
    hit[np.arange(n), np.arange(n)] = False
    return hit.any(1)


def cdist(XA, XB, metric, empty_mask=None, fill_val=None):
    XA = np.asarray(XA, np.float64)
    XB = np.asarray(XB, np.float64)
    if metric == 'cosine':
        with np.errstate(divide='ignore', invalid='ignore'):
            Y = 1. - (XA @ XB.T) / (np.linalg.norm(XA, axis=1)[:, None] * np.linalg.norm(XB, axis=1)[None, :])
    elif metric == 'euclidean':
        Y = np.sqrt(np.maximum(((XA[:, None, :] - XB[None, :, :]) ** 2).sum(-1), 0))
    else:
        raise ValueError(metric)
    if empty_mask is not None:
        Y = np.where(empty_mask, 1. if fill_val is None else fill_val, Y)
    return Y


def fuse_motion(cost, m_dist, w):
    out = (1. - w) * cost + w * (1. / CHI_SQ_INV_95) * m_dist
    out[m_dist > CHI_SQ_INV_95] = INF_COST
    return out


def gate_cost(cost, row_labels, col_labels, max_cost=None):
    cost = cost.copy()
    bad = np.asarray(row_labels)[:, None] != np.asarray(col_labels)[None, :]
    if max_cost is not None:
        bad |= cost > max_cost
    cost[bad] = INF_COST
    return cost


def lsa(cost):
    """Rectangular linear sum assignment, restating SciPy 1.18.1's shortest-augmenting-path solver
    (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp; Crouse 2016) including its scan order and tie
    rules.  Returns (rows, cols) like scipy.optimize.linear_sum_assignment."""
    cost = np.asarray(cost, np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transpose = nc < nr
    C = cost.T.copy() if transpose else cost
    nr, nc = C.shape
    u = np.zeros(nr)
    v = np.zeros(nc)
    path = np.full(nc, -1, np.int64)
    col4row = np.full(nr, -1, np.int64)
    row4col = np.full(nc, -1, np.int64)
    for cur in range(nr):
        remaining = list(range(nc - 1, -1, -1))
        SR = np.zeros(nr, bool)
        SC = np.zeros(nc, bool)
        spc = np.full(nc, np.inf)
        min_val = 0.0
        i = cur
        sink = -1
        while sink == -1:
            index = -1
            lowest = np.inf
            SR[i] = True
            for it, j in enumerate(remaining):
                r = min_val + C[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError('cost matrix is infeasible')
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[index] = remaining[-1]
            remaining.pop()
        u[cur] += min_val
        for r_ in range(nr):
            if SR[r_] and r_ != cur:
                u[r_] += min_val - spc[col4row[r_]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transpose:
        order = np.argsort(col4row, kind='stable')
        return col4row[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row


def numba_set_difference_order(n, removed):
    """Iteration order of `list(set(range(n)) - set(removed))` as compiled by Numba (matching.py:59-60).

    Numba's typed set is an open-addressing table (numba/cpython/setobj.py): hash(int) = int, 3 linear
    probes then i = (5 i + 1 + (perturb >>= 5)) & mask; `set(range(n))` preallocates the first power of two
    >= 2n (>= 16) and quadruples once if 2n hits the size; `a - b` copies, tombstones, then `downsize`s to the
    smallest power of two >= max(2*used, 16) when the table is >= 4x that, re-inserting survivors in slot
    order.  Without a downsize the order is ascending."""
    removed = set(int(x) for x in removed)
    keep = [v for v in range(n) if v not in removed]
    size = 16
    while size < 2 * n:
        size <<= 1
    if 2 * n >= size:
        size <<= 2
    min_entries = max(2 * len(keep), 16)
    if not (size >= 4 * min_entries and size > 16):
        return keep
    new_size = size
    while (new_size >> 1) >= min_entries:
        new_size >>= 1
    mask = new_size - 1
    table = [-1] * new_size
    for v in keep:
        i = v & mask
        perturb = v
        placed = False
        for _ in range(3):
            if table[i] == -1:
                table[i] = v
                placed = True
                break
            i = (i + 1) & mask
        while not placed:
            if table[i] == -1:
                table[i] = v
                placed = True
                break
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & mask
    return [v for v in table if v != -1]


def split_assignment(cost, row_ids, col_ids, m_rows, m_cols):
    """matching.py:57-70: unmatched in Numba-set order, then pairs demoted for cost >= INF in row order."""
    cost = np.asarray(cost)
    u_rows = [row_ids[r] for r in numba_set_difference_order(cost.shape[0], m_rows)]
    u_cols = [col_ids[c] for c in numba_set_difference_order(cost.shape[1], m_cols)]
    matches = []
    for r, c in zip(m_rows, m_cols):
        if cost[r, c] < INF_COST:
            matches.append((row_ids[r], col_ids[c]))
        else:
            u_rows.append(row_ids[r])
            u_cols.append(col_ids[c])
    return matches, u_rows, u_cols


def linear_assignment(cost, row_ids, col_ids):
    cost = np.asarray(cost, np.float64).reshape(len(row_ids), len(col_ids))
    m_rows, m_cols = lsa(cost)
    return split_assignment(cost, list(row_ids), list(col_ids), m_rows, m_cols)


def greedy_match(cost, row_ids, col_ids, max_cost):
    cost = np.asarray(cost, np.float64).reshape(len(row_ids), len(col_ids))
    rows = list(range(cost.shape[0]))
    cols = list(range(cost.shape[1]))
    matches = []
    while rows and cols:
        sub = cost[np.ix_(rows, cols)]
        k = int(np.argmin(sub))
        i, j = divmod(k, len(cols))
        if sub[i, j] <= max_cost:
            matches.append((row_ids[rows[i]], col_ids[cols[j]]))
            rows.pop(i)
            cols.pop(j)
        else:
            break
    return matches, [row_ids[r] for r in rows], [col_ids[c] for c in cols]
