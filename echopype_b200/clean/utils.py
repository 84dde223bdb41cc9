"""clean/utils.py:13-26 (extract_dB) and :380-401 (add_remove_background_noise_attrs) of the reference."""

import re


def extract_dB(dB_str: str) -> float:
    """Extract float value from decibel string in the form of 'NUMdB'."""
    if not isinstance(dB_str, str):
        raise TypeError(
            "Decibal input must be a string formatted as `NUMdB` or `NUMdb."
            f"Cannot be of type `{type(dB_str)}`."
        )
    match = re.search(r"^[-+]?\d+\.?\d*(?:dB|db)$", dB_str, flags=re.IGNORECASE)
    if match:
        return float(match.group(0)[:-2])
    raise ValueError("Decibal string must be formatted as 'NUMdB' or `NUMdb")


def noise_attrs(sv_type, lo, hi, ping_num, range_sample_num, SNR_threshold, noise_max):
    """Attributes of Sv_noise / Sv_corrected; lo/hi are the non-NaN extrema (NaN when the array is all NaN)."""
    return {
        "long_name": f"Volume backscattering strength, {sv_type} (Sv re 1 m-1)",
        "units": "dB",
        "actual_range": [round(float(lo), 2), round(float(hi), 2)],
        "noise_ping_num": ping_num,
        "noise_range_sample_num": range_sample_num,
        "SNR_threshold": SNR_threshold,
        "noise_max": noise_max,
    }
