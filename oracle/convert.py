"""TEST INFRASTRUCTURE ONLY - CPU restatement of the raw power ingest.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(echopype_b200/) never does.

Follows /root/reference/echopype/convert/parse_base.py:
  * :24      INDEX2POWER = 10.0 * np.log10(2.0) / 256.0
  * :686-730 ParseEK.pad_shorter_ping: pings shorter than the longest one are padded with NaN
  * :302     power = padded_arr.astype("float32") * INDEX2POWER   (a float64 product: INDEX2POWER is np.float64)
Pinned by tests/golden/ingest_vectors.npz, produced by executing those reference lines (make_golden_ingest.py).
"""

import numpy as np

INDEX2POWER = 10.0 * np.log10(2.0) / 256.0  # parse_base.py:24
PAD = np.int16(-32768)  # padding marker of the int16 ingest format (stands for the NaN of pad_shorter_ping)


def pad_shorter_ping(pings):
    """parse_base.py:686-730 for 1-D pings: (n_ping, max_len) float64, NaN beyond each ping's length."""
    lens = np.array([len(p) for p in pings])
    out = np.full((len(pings), int(lens.max())), np.nan)
    for i, p in enumerate(pings):
        out[i, : len(p)] = p
    return out


def power_from_counts(padded):
    """parse_base.py:302: counts (NaN-padded float array) -> power in dB, float64 as the reference computes it."""
    return padded.astype("float32").astype(np.float64) * INDEX2POWER


def pack_counts(pings):
    """The device ingest format: int16 [n_ping, max_len], PAD where pad_shorter_ping has NaN."""
    lens = np.array([len(p) for p in pings])
    out = np.full((len(pings), int(lens.max())), PAD, dtype=np.int16)
    for i, p in enumerate(pings):
        out[i, : len(p)] = p
    return out


def ingest_power_i16(counts):
    """What epb_ingest_power_i16 must return: float32(count * INDEX2POWER), PAD -> NaN."""
    c = np.asarray(counts)
    out = (c.astype(np.float64) * INDEX2POWER).astype(np.float32)
    out[c == PAD] = np.nan
    return out
