"""Host logic of the EK80 multi-filter_time paths (echopype_b200/calibrate/filter_time.py; reference
calibrate/api.py:95-197, calibrate_ek.py:25-52) on numpy-only Datasets: no GPU needed."""

import numpy as np
import pytest

from echopype_b200.calibrate import filter_time as ft
from echopype_b200.dataset import Dataset

T0 = np.datetime64("2024-05-01T00:00:00", "ns")


def _beam(P=10, chans=("WBT B", "WBT A"), first_valid=(0, 1)):
    pt = T0 + np.arange(P) * np.timedelta64(1, "s")
    tau = np.full((len(chans), P), 1e-3)
    for ci, f in enumerate(first_valid):
        tau[ci, :f] = np.nan
    return Dataset({"transmit_duration_nominal": (("channel", "ping_time"), tau)},
                   coords={"channel": np.array(chans, dtype=object), "ping_time": pt})


def _vend(beam, ft_idx=(0, 1, 6), chans=("WBT B", "WBT A")):
    pt = np.asarray(beam["ping_time"].values)
    C, F = len(chans), len(ft_idx)
    coeff = np.arange(C * F * 3, dtype=np.float64).reshape(C, F, 3)
    return Dataset({"PC_coeffs_real": (("channel", "filter_time", "PC_filter_n"), coeff),
                    "PC_deci_fac": (("channel", "filter_time"), np.arange(C * F).reshape(C, F)),
                    "impedance_transceiver": (("channel",), np.array([5400.0, 5000.0]))},
                   coords={"channel": np.array(chans, dtype=object), "filter_time": pt[list(ft_idx)]})


def test_first_valid_filter_time_and_collapse():
    beam = _beam()
    vend = _vend(beam)
    fv = ft.first_valid_filter_time_per_channel(beam)
    pt = np.asarray(beam["ping_time"].values)
    assert fv["WBT B"] == pt[0] and fv["WBT A"] == pt[1]
    out = ft.collapse_vend(vend, fv)
    assert "filter_time" not in out.sizes
    assert list(out["channel"].values) == ["WBT A", "WBT B"]  # the merge of the per-channel slices sorts the labels
    full = np.asarray(vend["PC_coeffs_real"].values)
    np.testing.assert_array_equal(out["PC_coeffs_real"].values, np.stack([full[1, 1], full[0, 0]]))
    np.testing.assert_array_equal(out["PC_deci_fac"].values, [np.asarray(vend["PC_deci_fac"].values)[1, 1], np.asarray(vend["PC_deci_fac"].values)[0, 0]])
    np.testing.assert_array_equal(out["impedance_transceiver"].values, [5000.0, 5400.0])
    with pytest.raises(KeyError):  # no filter set recorded at a channel's first valid ping (Dataset.sel would fail)
        ft.collapse_vend(_vend(beam, ft_idx=(0, 6)), fv)


def test_filter_pieces_follow_the_reference_loop():
    beam = _beam()
    vend = _vend(beam)
    pieces = ft.filter_pieces(beam, vend)
    # channels in sorted order ("WBT A" = index 1 first); intervals [filter time, next filter time - 1 ns]
    got = [(ci, p.tolist(), fi) for ci, p, fi in pieces]
    assert got == [
        (1, [1, 2, 3, 4, 5], 1), (1, [6, 7, 8, 9], 2),          # "WBT A": ping 0 has no valid pulse, filter time 0 is not its own
        (0, [0], 0), (0, [1, 2, 3, 4, 5], 1), (0, [6, 7, 8, 9], 2),
    ]


def _piece(chan, p_idx, value, tau_eff, P=10):
    pt = T0 + np.arange(P) * np.timedelta64(1, "s")
    n = len(p_idx)
    return Dataset({"Sv": (("channel", "ping_time", "range_sample"), np.full((1, n, 4), value)),
                    "tau_effective": (("channel",), np.array([tau_eff])),
                    "sound_speed": ((), np.float64(1500.0))},
                   coords={"channel": np.array([chan], dtype=object), "ping_time": pt[p_idx], "range_sample": np.arange(4)})


def test_merge_pieces_outer_join_and_conflicts():
    a = _piece("B", [0, 1, 2], -70.0, 1e-3)
    b = _piece("B", [5, 6], -60.0, 1e-3)
    c = _piece("A", [1, 2, 3], -50.0, 2e-3)
    out = ft.merge_pieces([a, b, c])
    assert list(out["channel"].values) == ["A", "B"]
    pt = T0 + np.array([0, 1, 2, 3, 5, 6]) * np.timedelta64(1, "s")
    np.testing.assert_array_equal(out["ping_time"].values, pt)  # union of the pieces' pings (ping 4 is in none)
    sv = out["Sv"].values
    assert sv.shape == (2, 6, 4)
    np.testing.assert_array_equal(sv[1, :3], -70.0)
    np.testing.assert_array_equal(sv[1, 4:], -60.0)
    assert np.isnan(sv[1, 3]).all() and np.isnan(sv[0, 0]).all() and np.isnan(sv[0, 4:]).all()
    np.testing.assert_array_equal(sv[0, 1:4], -50.0)
    np.testing.assert_array_equal(out["tau_effective"].values, [2e-3, 1e-3])
    assert float(out["sound_speed"].values) == 1500.0
    with pytest.raises(ValueError, match="conflicting values for variable 'tau_effective'"):
        ft.merge_pieces([a, _piece("B", [5, 6], -60.0, 3e-3)])
