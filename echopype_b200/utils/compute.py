"""dB <-> linear helpers (echopype/utils/compute.py:13-42) for host-side arrays.  On the device these are
fused into the kernels (csrc/epb_common.cuh db2lin / lin2db)."""

import numpy as np


def _log2lin(data):
    return 10 ** (np.asarray(data) / 10)


def _lin2log(data):
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10 * np.log10(data)
