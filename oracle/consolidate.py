"""CPU oracle (test infrastructure only): consolidate.add_depth, echopype/consolidate/api.py:163-221.

depth = transducer_depth + orientation * echo_range * echo_range_scaling, with transducer_depth a number or a
per-ping series aligned to ping_time (utils/align.py:5-61, interp "nearest" with extrapolation), and the scaling
cos(tilt) (number / per-ping series), the platform pitch-roll rotation (ek_depth_utils.py:55-75) or the normalised
beam z direction per channel (:78-120).
Pinned: reproduces the outputs of the reference's own add_depth / align_to_ping_time / ek_use_* functions
(tests/golden/make_golden_consolidate.py -> consolidate_vectors.npz, tests/test_reference_pinned_consolidate.py).
"""

import numpy as np
from scipy.interpolate import interp1d
from scipy.spatial.transform import Rotation


def align_nearest(values, t_ns, ping_ns):
    """utils/align.py:54-61: DataArray.interp(method="nearest", kwargs={"fill_value": "extrapolate"}) = scipy interp1d."""
    values = np.asarray(values, dtype=np.float64)
    if len(t_ns) == len(ping_ns) and np.array_equal(t_ns, ping_ns):
        return values
    if values.size == 1:
        return values.reshape(-1)[0] * np.ones(len(ping_ns))
    if values.size == 0:
        return np.full(len(ping_ns), np.nan)
    f = interp1d(np.asarray(t_ns, dtype=np.float64), values, kind="nearest", fill_value="extrapolate", assume_sorted=False)
    return f(np.asarray(ping_ns, dtype=np.float64))


def platform_angle_scaling(pitch_deg, roll_deg):
    """ek_depth_utils.py:66-70."""
    pitch_deg, roll_deg = np.atleast_1d(pitch_deg), np.atleast_1d(roll_deg)
    e = np.column_stack([np.zeros_like(pitch_deg), pitch_deg, roll_deg])
    return Rotation.from_euler("ZYX", e, degrees=True).as_matrix()[:, -1, -1]


def beam_angle_scaling(x, y, z):
    """ek_depth_utils.py:93-120."""
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    norm = np.sqrt(x**2 + y**2 + z**2)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(norm < 1e-8, np.nan, z / norm)


def add_depth(echo_range, transducer_depth=0.0, scaling=1.0, downward=True, per_channel=False):
    """consolidate/api.py:218-221.  echo_range (C,P,R) float64; transducer_depth scalar or (P,); scaling scalar, (P,) or
    (C,) when per_channel."""
    er = np.asarray(echo_range, dtype=np.float64)
    td = np.asarray(transducer_depth, dtype=np.float64)
    td = td[None, :, None] if td.ndim == 1 else td
    sc = np.asarray(scaling, dtype=np.float64)
    if sc.ndim == 1:
        sc = sc[:, None, None] if per_channel else sc[None, :, None]
    mult = 1 if downward else -1
    return td + (mult * er * sc)
