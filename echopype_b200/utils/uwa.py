"""Seawater acoustics used by the host-side parameter assembly (tiny arrays; stays on the CPU by design,
SURVEY.md 2 #7).  Same formulae and argument meaning as echopype/utils/uwa.py:8-189; coefficients are
tabulated instead of written inline."""

import numpy as np

_MACKENZIE = dict(c0=1448.96, t=(4.591, -5.304e-2, 2.374e-4), s=1.340, p=(1.630e-2, 1.675e-7), ts=-1.025e-2, tp3=-7.139e-13)


def calc_sound_speed(temperature=27, salinity=35, pressure=10, formula_source="Mackenzie"):
    """Sound speed [m/s]: "Mackenzie" (1981 nine-term) or "AZFP" (manufacturer's Matlab polynomial)."""
    T, dS, P = temperature, salinity - 35, pressure
    if formula_source == "Mackenzie":
        m = _MACKENZIE
        poly_t = m["c0"] + m["t"][0] * T + m["t"][1] * T**2 + m["t"][2] * T**3
        rest = m["s"] * dS + m["p"][0] * P + m["p"][1] * P**2
        cross = m["ts"] * T * dS + m["tp3"] * T * P**3
        return poly_t + rest + cross
    if formula_source == "AZFP":
        z, pk = T / 10, P / 1000
        return 1449.05 + z * (45.7 + z * (-5.21 + 0.23 * z)) + (1.333 + z * (-0.126 + z * 0.009)) * (salinity - 35.0) + pk * (16.3 + 0.18 * pk)
    raise UnboundLocalError("Unknown formula source")  # the reference falls through to an unbound local


def _absorption_fg(f_hz, T, S, P, pH, c):
    f = f_hz / 1000.0
    if c is None:
        c = 1412.0 + 3.21 * T + 1.19 * S + 0.0167 * P
    Tk = T + 273
    boric = (8.86 / c * 10 ** (0.78 * pH - 5), 1.0, 2.8 * np.sqrt(S / 35) * 10 ** (4 - 1245 / Tk))
    mgso4 = (
        21.44 * S / c * (1 + 0.025 * T),
        1.0 - 1.37e-4 * P + 6.2e-9 * P**2,
        8.17 * 10 ** (8 - 1990 / Tk) / (1 + 0.0018 * (S - 35)),
    )
    k = (4.937e-4, -2.59e-5, 9.11e-7, -1.5e-8) if np.all(T < 20) else (3.964e-4, -1.146e-5, 1.45e-7, -6.5e-10)
    A3 = k[0] + k[1] * T + k[2] * T**2 + k[3] * T**3
    P3 = 1.0 - 3.83e-5 * P + 4.9e-10 * P**2
    relax = lambda A, Pf, fr: A * Pf * fr * f**2 / (f**2 + fr**2)  # noqa: E731
    return (relax(*boric) + relax(*mgso4) + A3 * P3 * f**2) / 1000


def _absorption_am(f_hz, T, S, P, pH):
    f, D = f_hz / 1000, P / 1000
    f1 = 0.78 * np.sqrt(S / 35) * np.exp(T / 26)
    f2 = 42 * np.exp(T / 17)
    a1 = 0.106 * (f1 * (f**2)) / ((f1**2) + (f**2)) * np.exp((pH - 8) / 0.56)
    a2 = 0.52 * (1 + T / 43) * (S / 35) * (f2 * (f**2)) / ((f2**2) + (f**2)) * np.exp(-D / 6)
    a3 = 0.00049 * f**2 * np.exp(-(T / 27 + D))
    return (a1 + a2 + a3) / 1000


def _absorption_azfp(f_hz, T, S, P):
    Tk = T + 273.0
    f1 = 1320.0 * Tk * np.exp(-1700 / Tk)
    f2 = 1.55e7 * Tk * np.exp(-3052 / Tk)
    k = 1 + P / 10.0
    a = 8.95e-8 * (1 + T * (2.29e-2 - 5.08e-4 * T))
    b = (S / 35.0) * 4.88e-7 * (1 + 0.0134 * T) * (1 - 0.00103 * k + 3.7e-7 * k**2)
    c = 4.86e-13 * (1 + T * (-0.042 + T * (8.53e-4 - T * 6.23e-6))) * (1 + k * (-3.84e-4 + k * 7.57e-8))
    ff = f_hz**2
    if np.all(np.asarray(S) == 0):
        return c * ff
    return (a * f1 * ff) / (f1**2 + ff) + (b * f2 * ff) / (f2**2 + ff) + c * ff


def calc_absorption(frequency, temperature=27, salinity=35, pressure=10, pH=8.1, sound_speed=None, formula_source="AM"):
    """Absorption [dB/m]: "AM" (Ainslie & McColm 1998), "FG" (Francois & Garrison 1982) or "AZFP"."""
    if formula_source == "FG":
        return _absorption_fg(frequency, temperature, salinity, pressure, pH, sound_speed)
    if formula_source == "AM":
        return _absorption_am(frequency, temperature, salinity, pressure, pH)
    if formula_source == "AZFP":
        return _absorption_azfp(frequency, temperature, salinity, pressure)
    raise UnboundLocalError("Unknown formula source")
