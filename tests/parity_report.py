#!/usr/bin/env python
"""
Parity numbers of the round (needs a GPU): max |CUDA - reference| per output field
  * against what the UNMODIFIED reference wrote for the golden case (tests/golden/reference_glue.npz,
    oracle/make_golden_glue.py: step_03's pgw_for_era5 executed over oracle/xrlite.py), and
  * against the oracle on BASELINE configs[0] (201 x 281 x 137).
Prints one JSON object.

    python tests/parity_report.py > profiles/r1_parity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_timestep_gpu as T
    from cases import compare, make_case, run_oracle
    from pgw4era5_b200 import settings
    settings.i_debug = 0
    G, era, deltas, when = T._golden_case()
    out = {"golden_case": "5 x 6 columns, 137 levels, plev19, %s" % when.isoformat(), "vs_executed_reference": {}}
    fields = ("PS", "T", "QV", "U", "V", "T_SKIN", "T_SO", "FR_SEA_ICE")

    def diff(res, tag):
        d = {}
        for k in fields:
            g = res[k].detach().cpu().numpy().astype(np.float64).reshape(-1)
            r = G["pgw_%s_%s" % (tag, k)].astype(np.float64).reshape(-1)
            d[k] = float(np.nanmax(np.abs(g - r)))
        d["n_iter"] = [int(res["n_iter"]), int(G["pgw_%s_n_iter" % tag])]
        d["max_phi_error_per_iteration_diff"] = float(np.max(np.abs(np.array(res["phi_max_errors"]) - G["pgw_%s_errs" % tag])))
        return d

    res, _ = T._apply(era, deltas, when=when)
    out["vs_executed_reference"]["PS_FIS_double (default64)"] = diff(res, "default64")
    out["vs_executed_reference"]["all_float32_file (default)"] = diff(res, "default")
    for tag, name, value in (("tight64", "thresh_phi_ref_max_error", 1e-3), ("pref_none64", "p_ref_inp", None),
                             ("reinterp64", "i_reinterp", 1)):
        old = getattr(settings, name)
        setattr(settings, name, value)
        try:
            res, _ = T._apply(era, deltas, when=when)
        finally:
            setattr(settings, name, old)
        out["vs_executed_reference"][tag] = diff(res, tag)
    era, deltas = make_case(201, 281, 1)
    ref = run_oracle(era, deltas)
    res, _ = T._apply(era, deltas)
    out["vs_oracle_config1_201x281"] = dict(compare(res, ref), n_iter=[int(res["n_iter"]), int(ref["n_iter"])])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
