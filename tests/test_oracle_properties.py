"""Known-answer and property tests of the CPU oracle and the
pins against scipy's interp1d (the third-party arithmetic behind xarray .interp)."""
from datetime import datetime

import numpy as np
import pytest

from cases import ERA_DATE, make_case, run_oracle
from oracle import pgw_oracle as O
from pgw4era5_b200 import synthetic as S


def test_interp1d_axis_matches_scipy():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(-90, 90, 17))
    y = rng.normal(size=(4, 17, 5))
    xn = np.concatenate([rng.uniform(x[0], x[-1], 20), x[[0, 5, 16]], [x[0] - 1, x[-1] + 1]])
    ref = interp1d(x, y, axis=1, kind="linear", bounds_error=False, fill_value=np.nan)(xn)
    got = O._interp1d_axis(x, y, xn, axis=1)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)


def test_time_interp_matches_scipy():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(4)
    lo, hi = rng.normal(size=(3, 4)).astype(np.float32), rng.normal(size=(3, 4)).astype(np.float32)
    x_hi, x_new = 2678400e9, 1447200e9
    ref = interp1d(np.array([0.0, x_hi]), np.stack([lo, hi]), axis=0)(x_new)
    got = O.interp1d_linear_2pt(lo.astype(np.float64), hi.astype(np.float64), x_hi, x_new)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-15)


def test_load_delta_time_semantics():
    stamps = S.monthly_stamps()
    data = np.arange(12, dtype=np.float64).reshape(12, 1, 1) * np.ones((12, 2, 3))
    d = dict(time=stamps, plev=None, data=data)
    np.testing.assert_array_equal(O.load_delta(d, datetime(2006, 3, 16, 12))[0], data[2])     # exact hit
    jan1 = O.load_delta(d, datetime(2006, 1, 1, 0))[0]                                         # Dec(-1) .. Jan
    assert 0 < jan1[0, 0] < 11 and abs(jan1[0, 0] - (11 + (0 - 11) * 1339200 / 2678400)) < 1e-12
    dec31 = O.load_delta(d, datetime(2006, 12, 31, 18))[0]                                     # Dec .. Jan(+1)
    assert abs(dec31[0, 0] - (11 + (0 - 11) * 1317600 / 2678400)) < 1e-12
    assert O.load_delta(d, None).shape == (12, 2, 3)


def test_integ_geopot_isothermal_dry():
    """Dry isothermal column: phi(p_ref) = FIS + Rd T ln(ps/p_ref), independent of the levels."""
    ak, bk = S.hybrid_coefficients()
    ps = np.array([[[101325.0, 95000.0, 70000.0]]])
    pa_hl = ak[None, :, None, None] + ps[:, None] * bk[None, :, None, None]
    ta = np.full((1, 137, 1, 3), 250.0)
    phi = O.integ_geopot(pa_hl, np.full((1, 1, 3), 100.0), ta, np.zeros_like(ta), 30000.0)
    np.testing.assert_allclose(phi, 100.0 + O.CON_RD * 250.0 * np.log(ps / 30000.0), rtol=1e-13)
    with pytest.raises(ValueError, match="below the surface"):
        O.integ_geopot(pa_hl, np.zeros((1, 1, 3)), ta, np.zeros_like(ta), 80000.0)


def test_zero_deltas_one_iteration_identity():
    era, deltas = make_case(6, 9, 21)
    for v in deltas.values():
        v["data"].zero_()
    deltas["ps_hist"]["data"] += era["PS"]
    ref = run_oracle(era, deltas)
    # not ~0: the reference multiplies the float32 T and QV of the ERA state in float32 (functions.py:144),
    # the float64 PGW state in float64 (reproduced by the oracle, pinned in test_oracle_glue_golden.py)
    assert ref["n_iter"] == 1 and ref["phi_max_errors"][0] < 2e-2
    e = S.to_numpy(era)
    np.testing.assert_array_equal(ref["PS"], e["PS"].astype(np.float64))
    np.testing.assert_array_equal(ref["T"], e["T"].astype(np.float64))
    np.testing.assert_array_equal(ref["U"], e["U"].astype(np.float64))
    np.testing.assert_allclose(ref["QV"], e["QV"], rtol=1e-12, atol=0)      # rh -> q round trip


def test_hydrostatic_consistency_uniform_warming():
    """Uniform warming dT with the matching hypsometric zg delta at p_ref leaves ps unchanged (item 9)."""
    era, deltas = make_case(5, 7, 22)
    e = S.to_numpy(era)
    dT = 2.0
    for name in ("hur", "ua", "va", "hurs", "siconc"):
        deltas[name]["data"].zero_()
    for name in ("ta", "tas", "ts", "tos"):
        deltas[name]["data"].fill_(dT)
    deltas["ps_hist"]["data"].zero_()
    deltas["ps_hist"]["data"] += era["PS"]
    zero = run_oracle(era, {k: dict(v, data=v["data"] * 0 if k != "ps_hist" else v["data"]) for k, v in deltas.items()})
    # zg delta at p_ref that makes dps = 0: (phi_pgw(ps) - phi_era(ps)) / g, evaluated by the oracle itself
    ak, bk = e["ak"], e["bk"]
    PS = e["PS"].astype(np.float64)
    pa_hl = ak[None, :, None, None] + PS[:, None] * bk[None, :, None, None]
    akm, bkm = e["akm"], e["bkm"]
    pa = akm[None, :, None, None] + PS[:, None] * bkm[None, :, None, None]
    T = e["T"].astype(np.float64); QV = e["QV"].astype(np.float64)
    rh = O.specific_to_relative_humidity(QV, pa, T)
    q2 = O.relative_to_specific_humidity(rh, pa, T + dT)
    dphi = O.integ_geopot(pa_hl, e["FIS"], T + dT, q2, 30000.0) - O.integ_geopot(pa_hl, e["FIS"], T, QV, 30000.0)
    k300 = int(np.nonzero(S.PLEV19 == 30000.0)[0][0])
    deltas["zg"]["data"].zero_()
    import torch
    deltas["zg"]["data"][:, k300] = torch.as_tensor(dphi[0] / O.CON_G, dtype=torch.float32)
    ref = run_oracle(era, deltas)
    assert ref["n_iter"] == 1
    assert np.max(np.abs(ref["PS"] - PS)) == 0.0
    assert zero["n_iter"] == 1


def test_oracle_raises_like_reference():
    era, deltas = make_case(5, 7, 23)
    with pytest.raises(ValueError, match="top pressure"):
        run_oracle(era, deltas, ignore_top_pressure_error=False)
    with pytest.raises(ValueError, match="did not converge"):
        run_oracle(era, deltas, max_n_iter=3)
    with pytest.raises(ValueError, match="below the surface"):
        run_oracle(era, deltas, p_ref_inp=100000)
    bad = {k: dict(v) for k, v in deltas.items()}
    bad["ps_hist"] = dict(deltas["ps_hist"], data=deltas["ps_hist"]["data"] * 0 + 50.0)
    with pytest.raises(ValueError):
        run_oracle(era, bad)


def test_oracle_variants_i_reinterp_and_local_p_ref():
    """step_03:202-251,:330-343 restated in the oracle: sanity of the two non-default settings."""
    import numpy as np
    from cases import make_case, run_oracle
    era, deltas = make_case(5, 8, 3)
    base = run_oracle(era, deltas)
    same = run_oracle(era, deltas, i_reinterp=0, p_ref_inp=30000)
    assert same["n_iter"] == base["n_iter"] and np.array_equal(same["PS"], base["PS"])
    re = run_oracle(era, deltas, i_reinterp=1)
    loc = run_oracle(era, deltas, p_ref_inp=None)
    # the first iteration does not depend on i_reinterp (pa_pgw == pa_era): same first error
    assert abs(re["phi_max_errors"][0] - base["phi_max_errors"][0]) < 1e-9
    plev = np.asarray(deltas["zg"]["plev"], dtype=np.float64)
    assert np.all(np.isin(loc["p_ref"], plev))
    ps = np.asarray(era["PS"], dtype=np.float64)
    assert np.all(loc["p_ref"] < 0.95 * np.minimum(ps, loc["PS"]) + 1e-9)
    for r in (re, loc):
        assert r["phi_max_errors"][-1] <= 0.15 < r["phi_max_errors"][-2]
        assert np.all(np.isfinite(r["PS"])) and float(np.abs(r["PS"] - base["PS"]).max()) < 200.0
