#!/usr/bin/python
# -*- coding: utf-8 -*-
"""
Interpolate CFday (GCM model-level) output to pressure levels: drop-in for the reference's
``step_01_extract_deltas/CFday_interp_to_plev.py``.  Same positional arguments
(``var_names experiment``) and file-name pattern; the author-specific directories and the list of
5-year chunks of the reference (:38-69) are options here.  The interpolation itself
(:113-127, interp_logp_4d with constant extrapolation on ascending pressure) runs on the GPU
through ``functions.interp_logp_4d`` -> ``pgw_interp_logp_f32/f64``.

    python -m pgw4era5_b200.step_01_extract_deltas.CFday_interp_to_plev ta,hur ssp585 \\
           --inp_dir subdomain --out_base_dir interp_plev --target_p_file CFday_target_p_MPI-ESM1-2-HR.dat
"""
import argparse
import os
from pathlib import Path

import numpy as np

from .. import ncio
from ..functions import interp_logp_4d
from ..settings import LAT_GCM, LEV_GCM, LON_GCM, PLEV_GCM, TIME_GCM

TIMES = {                                                     # CFday_interp_to_plev.py:52-69
    'ssp585': ['20700101-20741231', '20750101-20791231', '20800101-20841231',
               '20850101-20891231', '20900101-20941231', '20950101-20991231'],
    'historical': ['19850101-19891231', '19900101-19941231', '19950101-19991231',
                   '20000101-20041231', '20050101-20091231', '20100101-20141231'],
}


def interp_dataset_to_plev(ds, var_name, targ_plev):
    """CFday_interp_to_plev.py:92-159 on an ``ncio.Dataset`` with ap, b [lev], ps [time, lat, lon] and
    ``var_name`` [time, lev, lat, lon]: returns a new Dataset with the variable on ``targ_plev``
    (stored with descending pressure like the reference, :131-134)."""
    var = ds[var_name]
    if var.dims != (TIME_GCM, LEV_GCM, LAT_GCM, LON_GCM):
        raise ValueError("%s must be (%s, %s, %s, %s)" % (var_name, TIME_GCM, LEV_GCM, LAT_GCM, LON_GCM))
    # sort pressure ascending (:95): the lev axis is reversed
    data = np.asarray(var.data)[:, ::-1]
    ap = np.asarray(ds['ap'].data, dtype=np.float64)[::-1]
    b = np.asarray(ds['b'].data, dtype=np.float64)[::-1]
    ps = np.asarray(ds['ps'].data, dtype=np.float64)
    # pressure on full levels (:97-98)
    source_P = ap[None, :, None, None] + b[None, :, None, None] * ps[:, None]
    targ_plev = np.sort(np.asarray(targ_plev, dtype=np.float64))             # :113
    nt, _, ny, nx = data.shape
    targ_P = np.broadcast_to(targ_plev[None, :, None, None], (nt, len(targ_plev), ny, nx))   # :117-121
    var_out = interp_logp_4d(np.ascontiguousarray(data, dtype=np.float64), source_P,
                             np.ascontiguousarray(targ_P), extrapolate='constant',
                             time_key=TIME_GCM, lat_key=LAT_GCM, lon_key=LON_GCM)             # :124-127
    out = ncio.Dataset(attrs=ds.attrs)
    for name in (TIME_GCM, LAT_GCM, LON_GCM):
        if name in ds:
            out[name] = ds[name]
    out[PLEV_GCM] = ncio.Variable((PLEV_GCM,), targ_plev[::-1].copy())       # descending (:131-134)
    out[var_name] = ncio.Variable((TIME_GCM, PLEV_GCM, LAT_GCM, LON_GCM),
                                  np.ascontiguousarray(np.asarray(var_out)[:, ::-1]).astype(var.data.dtype),
                                  var.attrs)
    return out


def main(argv=None):
    parser = argparse.ArgumentParser(description='Interpolate CFday output to pressure levels')
    parser.add_argument('var_names', type=str)
    parser.add_argument('experiment', type=str)
    parser.add_argument('--inp_dir', type=str, default='subdomain')
    parser.add_argument('--out_base_dir', type=str, default='interp_plev')
    parser.add_argument('--model_name', type=str, default='MPI-ESM1-2-HR')
    parser.add_argument('--target_p_file', type=str, default='CFday_target_p_MPI-ESM1-2-HR.dat')
    parser.add_argument('--times', type=str, default=None,
                        help='comma separated YYYYMMDD-YYYYMMDD chunks (default: the reference\'s list)')
    args = parser.parse_args(argv)
    chunks = args.times.split(',') if args.times else TIMES[args.experiment]
    targ_plev = np.loadtxt(args.target_p_file)
    for var_name in args.var_names.split(','):
        for time_ind, chunk in enumerate(chunks):
            print(time_ind)
            file_name = '{}_CFday_{}_{}_r1i1p1f1_gn_{}.nc'.format(var_name, args.model_name, args.experiment, chunk)
            out_dir = os.path.join(args.out_base_dir, args.model_name)
            Path(out_dir).mkdir(parents=True, exist_ok=True)
            inp_file_path, out_file_path = os.path.join(args.inp_dir, file_name), os.path.join(out_dir, file_name)
            print('Process input file: \n{}\nto output file: \n{}'.format(inp_file_path, out_file_path))
            ds = ncio.open_dataset(inp_file_path)
            interp_dataset_to_plev(ds, var_name, targ_plev).to_netcdf(out_file_path)


if __name__ == '__main__':
    main()
