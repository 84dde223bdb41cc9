"""CPU oracle (test infrastructure only): mask.frequency_differencing and mask.apply_mask.

frequency_differencing, echopype/mask/api.py:593-608: mask = (Sv[chanA] - Sv[chanB]) <operator> diff, NaN -> False.
apply_mask, echopype/mask/api.py:395-438: masks broadcast over channel, combined with logical AND, NaN mask entries
count as False, result = where(mask, var, fill_value).
frequency_differencing is pinned to the outputs of the reference's own function (tests/golden/make_golden_mask.py ->
freqdiff_vectors.npz, tests/test_reference_pinned_consolidate.py), and so is apply_mask (lifted with its helpers, four mask /
fill combinations).
"""

import operator as _op

import numpy as np

STR2OPS = {">": _op.gt, "<": _op.lt, "<=": _op.le, ">=": _op.ge, "==": _op.eq}


def frequency_differencing(Sv, chanA_idx, chanB_idx, operator, diff):
    Sv = np.asarray(Sv, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return np.where(STR2OPS[operator](Sv[chanA_idx] - Sv[chanB_idx], diff), True, False)


def apply_mask(var, masks, fill_value=np.nan):
    var = np.asarray(var, dtype=np.float64)
    final = np.ones(var.shape, dtype=bool)
    for m in masks:
        m = np.asarray(m)
        if m.dtype.kind == "f":
            m = np.where(np.isnan(m), False, m)
        m = m.astype(bool)
        final &= np.broadcast_to(m if m.ndim == var.ndim else m[None], var.shape)
    return np.where(final, var, fill_value)
