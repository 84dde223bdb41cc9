"""Argument parsing / validation of mask.frequency_differencing (echopype/mask/freq_diff.py:7-146): same patterns,
same error types and messages."""

import re

import numpy as np


def _parse_freq_diff_eq(freqABEq=None, chanABEq=None):
    if (freqABEq is None) and (chanABEq is None):
        raise ValueError("Either freqAB or chanAB must be given!")
    elif (freqABEq is not None) and (chanABEq is not None):
        raise ValueError("Only one of freqAB or chanAB should be given, but not both!")
    operatorPattern = r"\s*(?P<cmp>\S*?)\s*"
    rhsPattern = r"(?P<db>\d*\.?\d+)\s*dB"
    if freqABEq is not None:
        freqAPattern = r"(?P<freqA>\d*\.?\d+)\s*(?P<unitA>\w?)Hz"
        freqBPattern = r"(?P<freqB>\d*\.?\d+)\s*(?P<unitB>\w?)Hz"
        m = re.compile(freqAPattern + r"\s*-\s*" + freqBPattern + operatorPattern + rhsPattern).match(freqABEq)
        if m is None:
            raise TypeError("Invalid freqAB Equation!")
        operator = m["cmp"]
        if operator not in [">", "<", "<=", ">=", "=="]:
            raise ValueError("Invalid operator!")
        mult = {"": 1, "k": 1e3, "M": 1e6, "G": 1e9}
        freqAB = [float(m["freqA"]) * mult[m["unitA"]], float(m["freqB"]) * mult[m["unitB"]]]
        if len(set(freqAB)) != 2:
            raise ValueError("freqAB must be a list of length 2 with unique elements!")
        return [freqAB, None, operator, float(m["db"])]
    chanAPattern = r"(?P<chanA>\".+\")\s*"
    chanBPattern = r"(?P<chanB>\".+\")\s*"
    m = re.compile(chanAPattern + r"\s*-\s*" + chanBPattern + operatorPattern + rhsPattern).match(chanABEq)
    if m is None:
        raise TypeError("Invalid chanAB Equation!")
    operator = m["cmp"]
    if operator not in [">", "<", "<=", ">=", "=="]:
        raise ValueError("Invalid operator!")
    chanAB = [m["chanA"][1:-1], m["chanB"][1:-1]]
    if len(set(chanAB)) != 2:
        raise ValueError("chanAB must be a list of length 2 with unique elements!")
    return [None, chanAB, operator, float(m["db"])]


def _check_freq_diff_source_Sv(source_Sv, freqAB=None, chanAB=None):
    if "channel" not in source_Sv.coords:
        raise ValueError("The Dataset defined by source_Sv must have channel as a coordinate!")
    elif "frequency_nominal" not in source_Sv.variables:
        raise ValueError("The Dataset defined by source_Sv must have frequency_nominal as a variable!")
    chan = [str(c) for c in np.asarray(source_Sv["channel"].values).tolist()]
    if chanAB is not None:
        if len(set(chan)) < len(chan):
            raise ValueError("The provided source_Sv contains repeated channel values, this is not allowed!")
        if not all(c in chan for c in chanAB):
            raise ValueError("The provided list input chanAB contains values that are not in the channel coordinate!")
    if freqAB is not None:
        f = np.asarray(source_Sv["frequency_nominal"].values, dtype=np.float64).tolist()
        if len(set(f)) < len(f):
            raise ValueError("The provided source_Sv contains repeated frequency_nominal values, this is not allowed!")
        if not all(x in f for x in freqAB):
            raise ValueError("The provided list input freqAB contains values that are not in the frequency_nominal variable!")
