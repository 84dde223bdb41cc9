"""The synthetic calibration cases shared by ``make_golden_calibrate.py`` (which runs the REFERENCE's own
calibration classes on them) and by the tests that compare the oracle and the CUDA path with the stored
reference outputs (tests/golden/calibrate_vectors.npz)."""

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = os.path.join(HERE, "calibrate_vectors.npz")

# key -> (synth maker, maker kwargs, calibrate kwargs)
CASES = {
    "ek60_plain": ("ek60", dict(C=3, P=12, R=96, seed=11, nan_tail=0.3), {}),
    "ek60_tv": ("ek60", dict(C=4, P=16, R=80, seed=12, nan_tail=0.3, time_varying=True), {}),
    "ek60_env": ("ek60", dict(C=3, P=12, R=96, seed=11, nan_tail=0.3),
                 dict(env_params={"temperature": 8.5, "salinity": 31.0, "pressure": 40.0, "pH": 8.0})),
    "ek60_cal": ("ek60", dict(C=3, P=12, R=96, seed=11, nan_tail=0.3),
                 dict(cal_params={"gain_correction": [25.0, 26.5, 27.5], "equivalent_beam_angle": -20.0})),
    "noise": ("ek60", dict(C=3, P=23, R=200, seed=13, nan_tail=0.2), {}),
    "ek80_cw_power": ("ek80", dict(C=3, P=10, R=96, seed=21, mode="CW", encode="power", gpt_channel=1, nan_tail=0.3),
                      dict(waveform_mode="CW", encode_mode="power")),
    "ek80_cw_complex": ("ek80", dict(C=2, P=8, R=96, B=4, seed=22, mode="CW", encode="complex", nan_tail=0.3),
                        dict(waveform_mode="CW", encode_mode="complex")),
    "ek80_bb": ("ek80", dict(C=2, P=5, R=640, B=4, seed=23, mode="BB", encode="complex", nan_tail=0.4),
                dict(waveform_mode="BB", encode_mode="complex")),
    "ek80_bb_drop": ("ek80", dict(C=2, P=5, R=640, B=4, seed=23, mode="BB", encode="complex", nan_tail=0.4),
                     dict(waveform_mode="BB", encode_mode="complex", drop_last_hanning_zero=True)),
    "azfp": ("azfp", dict(C=4, P=9, R=120, seed=31), dict(env_params={"salinity": 29.6, "pressure": 60.0})),
}
CAL_TYPES = {"ek60_env": ("Sv",), "ek60_cal": ("Sv",), "ek80_bb_drop": ("Sv",), "noise": ("Sv",)}
# remove_background_noise settings run by the reference on the "noise" case: (ping_num, range_sample_num, max, SNR)
NOISE_ARGS = {"a": (5, 30, None, "3.0dB"), "b": (10, 20, "-125.0dB", "3.0dB"), "c": (4, 25, "-140dB", "6.5dB")}
# cases whose backscatter is stored in the vector file under <key>__in__*; the others share one of these
INPUT_OF = {"ek60_env": "ek60_plain", "ek60_cal": "ek60_plain", "ek80_bb_drop": "ek80_bb"}


def cal_types(key):
    return CAL_TYPES.get(key, ("Sv", "TS"))


def raw_backscatter(key):
    """the generator-side inputs (numpy RNG of echopype_b200.synth) incl. the hand-placed beam gaps"""
    from echopype_b200 import synth

    maker, kw, _ = CASES[key]
    ed = getattr(synth, "make_" + maker)(**kw)
    b = ed["Sonar/Beam_group1"]
    re_ = np.array(b["backscatter_r"].values)
    if "backscatter_i" not in b:
        return re_
    im_ = np.array(b["backscatter_i"].values)
    if key == "ek80_cw_complex":  # one sector missing on some samples: the beam mean skips it (xarray skipna)
        re_[0, 2, 10:20, 3] = np.nan
        im_[0, 2, 10:20, 3] = np.nan
    return re_, im_


def build(key, vectors=None):
    """EchoData of a case.  With ``vectors`` (the loaded npz) the stored backscatter is used, so the
    case does not depend on the numpy RNG stream of the machine running the test."""
    from echopype_b200 import synth

    maker, kw, _ = CASES[key]
    src = INPUT_OF.get(key, key)
    if vectors is None:
        bs = raw_backscatter(src)
    elif f"{src}__in__backscatter_i" in vectors:
        bs = (vectors[f"{src}__in__backscatter_r"], vectors[f"{src}__in__backscatter_i"])
    else:
        bs = vectors[f"{src}__in__backscatter_r"]
    return getattr(synth, "make_" + maker)(backscatter=bs, **kw)


def load():
    return np.load(VECTORS, allow_pickle=False)
