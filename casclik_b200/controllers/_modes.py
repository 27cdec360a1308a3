"""Mode enumeration of the set-based pseudo-inverse controller.

Contract (reference casclik/controllers/pseudo_inverse.py:107-130, pinned by
tests/golden/activation_maps.json which was produced by running the reference's own function):
all 2^S activation patterns of the S SetConstraints, entry k = 1 iff the k-th SetConstraint in
priority order is active, ordered by number of active sets and, within equal counts, by the
integer whose bit k is entry k.
"""


def activation_map(n_sets):
    if n_sets == 0:
        return []
    patterns = [[(code >> k) & 1 for k in range(n_sets)] for code in range(1 << n_sets)]
    patterns.sort(key=sum)   # stable: ascending code within equal popcount
    return patterns


def mode_masks(n_sets):
    return [sum(bit << k for k, bit in enumerate(row)) for row in activation_map(n_sets)] or [0]
