"""Host-side replay of the one Numba container behaviour the reference's results depend on.

`_get_assignment_matches` (fastmot/utils/matching.py:57-70) builds its unmatched lists with
`list(set(range(n)) - set(matched))` inside an @njit function, so the order in which unmatched
detections reach the next cascade stage — and therefore the order in which new track IDs are handed out
(fastmot/tracker.py:286-293) — is the slot order of Numba's open-addressing set.  This module computes
that order for non-negative ints without Numba.
"""
import numpy as np

_MINSIZE = 16


def set_difference_order(n, removed):
    """Order of `list(set(range(n)) - set(removed))` under Numba's typed set."""
    mask_keep = np.ones(n, bool)
    if len(removed):
        mask_keep[np.asarray(removed, np.int64)] = False
    keep = np.nonzero(mask_keep)[0]
    size = _MINSIZE
    while size < 2 * n:
        size <<= 1
    if 2 * n >= size:          # the add that fills half the table quadruples it
        size <<= 2
    min_entries = max(2 * len(keep), _MINSIZE)
    if not (size >= 4 * min_entries and size > _MINSIZE):
        return keep.tolist()   # no shrink: slot == value, ascending
    new_size = size
    while (new_size >> 1) >= min_entries:
        new_size >>= 1
    m = new_size - 1
    table = [-1] * new_size
    for v in keep.tolist():    # survivors are re-inserted in old slot (= ascending) order
        i = v & m
        for _ in range(3):     # three linear probes ...
            if table[i] < 0:
                break
            i = (i + 1) & m
        else:
            perturb = v        # ... then the perturbed sequence
            while table[i] >= 0:
                perturb >>= 5
                i = (i * 5 + 1 + perturb) & m
        table[i] = v
    return [v for v in table if v >= 0]
