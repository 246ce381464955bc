"""Parity of the fused CUDA timestep (through the C ABI) against the fp64 oracle."""
import numpy as np
import pytest
import torch

from cases import ERA_DATE, TOL, compare, make_case, run_oracle

pytestmark = pytest.mark.gpu


def _engine(era, deltas):
    from pgw4era5_b200.engine import DeltaSet, PGWEngine
    ds = DeltaSet(deltas, device="cuda")
    return PGWEngine(era["ak"], era["bk"], ds, soil1=era["soil1"])


def _dev(era):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in era.items()}


@pytest.mark.parametrize("ny,nx,seed", [(24, 40, 1), (33, 129, 2), (7, 31, 3)])
def test_timestep_matches_oracle(ny, nx, seed):
    era, deltas = make_case(ny, nx, seed)
    ref = run_oracle(era, deltas)
    eng = _engine(era, deltas)
    res = eng.apply(_dev(era), ERA_DATE, ignore_top_pressure_error=True)
    assert res["n_iter"] == ref["n_iter"]
    np.testing.assert_allclose(res["phi_max_errors"], ref["phi_max_errors"], rtol=0, atol=1e-3)
    errs = compare(res, ref)
    for name, e in errs.items():
        assert e <= TOL[name], (name, e, errs)


@pytest.mark.parametrize("k_spec", [3, 5, 9, 19])
def test_speculation_paths_agree(k_spec):
    """rerun (k_spec too small), exact hit and rewrite (k_spec too large) give the same fields."""
    era, deltas = make_case(24, 40, 1)
    ref = run_oracle(era, deltas)
    eng = _engine(era, deltas)
    res = eng.apply(_dev(era), ERA_DATE, ignore_top_pressure_error=True, k_spec=k_spec)
    assert res["n_iter"] == ref["n_iter"]
    errs = compare(res, ref)
    for name, e in errs.items():
        assert e <= TOL[name], (name, e, errs)
