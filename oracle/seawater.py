"""Oracle: seawater sound speed / absorption (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows echopype/utils/uwa.py:8-53 (sound speed) and :56-189 (absorption).
"""

import numpy as np


def sound_speed(temperature=27, salinity=35, pressure=10, formula_source="Mackenzie"):
    """uwa.py:8-53.  Mackenzie (1981) nine-term equation, or the AZFP Matlab polynomial."""
    T, S, P = temperature, salinity, pressure
    if formula_source == "Mackenzie":
        c = 1448.96 + 4.591 * T - 5.304e-2 * T**2 + 2.374e-4 * T**3
        c = c + (1.340 * (S - 35) + 1.630e-2 * P + 1.675e-7 * P**2)
        c = c + (-1.025e-2 * T * (S - 35) - 7.139e-13 * T * P**3)
        return c
    if formula_source == "AZFP":
        z = T / 10
        return (
            1449.05
            + z * (45.7 + z * (-5.21 + 0.23 * z))
            + (1.333 + z * (-0.126 + z * 0.009)) * (S - 35.0)
            + (P / 1000) * (16.3 + 0.18 * (P / 1000))
        )
    # uwa.py:51-52 constructs (does not raise) a ValueError, then hits an unbound local
    raise UnboundLocalError("Unknown formula source")


def absorption(
    frequency, temperature=27, salinity=35, pressure=10, pH=8.1, sound_speed=None, formula_source="AM"
):
    """uwa.py:56-189.  Returns dB/m."""
    T, S, P = temperature, salinity, pressure
    if formula_source == "FG":  # Francois & Garrison 1982, uwa.py:109-140
        f = frequency / 1000.0
        c = (1412.0 + 3.21 * T + 1.19 * S + 0.0167 * P) if sound_speed is None else sound_speed
        A1 = 8.86 / c * 10 ** (0.78 * pH - 5)
        f1 = 2.8 * np.sqrt(S / 35) * 10 ** (4 - 1245 / (T + 273))
        A2 = 21.44 * S / c * (1 + 0.025 * T)
        P2 = 1.0 - 1.37e-4 * P + 6.2e-9 * P**2
        f2 = 8.17 * 10 ** (8 - 1990 / (T + 273)) / (1 + 0.0018 * (S - 35))
        P3 = 1.0 - 3.83e-5 * P + 4.9e-10 * P**2
        if np.all(T < 20):
            A3 = 4.937e-4 - 2.59e-5 * T + 9.11e-7 * T**2 - 1.5e-8 * T**3
        else:
            A3 = 3.964e-4 - 1.146e-5 * T + 1.45e-7 * T**2 - 6.5e-10 * T**3
        a = (
            A1 * 1.0 * f1 * f**2 / (f**2 + f1**2)
            + A2 * P2 * f2 * f**2 / (f**2 + f2**2)
            + A3 * P3 * f**2
        )
        return a / 1000
    if formula_source == "AM":  # Ainslie & McColm 1998, uwa.py:142-157
        fk = frequency / 1000
        D = P / 1000
        f1 = 0.78 * np.sqrt(S / 35) * np.exp(T / 26)
        f2 = 42 * np.exp(T / 17)
        a1 = 0.106 * (f1 * (fk**2)) / ((f1**2) + (fk**2)) * np.exp((pH - 8) / 0.56)
        a2 = 0.52 * (1 + T / 43) * (S / 35) * (f2 * (fk**2)) / ((f2**2) + (fk**2)) * np.exp(-D / 6)
        a3 = 0.00049 * fk**2 * np.exp(-(T / 27 + D))
        return (a1 + a2 + a3) / 1000
    if formula_source == "AZFP":  # uwa.py:159-185
        Tk = T + 273.0
        f1 = 1320.0 * Tk * np.exp(-1700 / Tk)
        f2 = 1.55e7 * Tk * np.exp(-3052 / Tk)
        k = 1 + P / 10.0
        a = 8.95e-8 * (1 + T * (2.29e-2 - 5.08e-4 * T))
        b = (S / 35.0) * 4.88e-7 * (1 + 0.0134 * T) * (1 - 0.00103 * k + 3.7e-7 * k**2)
        c = (
            4.86e-13
            * (1 + T * (-0.042 + T * (8.53e-4 - T * 6.23e-6)))
            * (1 + k * (-3.84e-4 + k * 7.57e-8))
        )
        if np.all(np.asarray(S) == 0):
            return c * frequency**2
        return (
            (a * f1 * frequency**2) / (f1**2 + frequency**2)
            + (b * f2 * frequency**2) / (f2**2 + frequency**2)
            + c * frequency**2
        )
    raise UnboundLocalError("Unknown formula source")
