#!/usr/bin/python
# -*- coding: utf-8 -*-
"""
PGW for ERA5, preprocessing of climate deltas: drop-in for the reference's
``step_02_preproc_deltas.py`` (same positional ``processing_step`` and ``-i -o -e -v``).
Smoothing (harmonic annual-cycle filter) and bilinear regridding run as CUDA kernels.

    python -m pgw4era5_b200.step_02_preproc_deltas regridding -i gcm -o out -e era5.nc -v ta,hur
"""
import argparse
import os
from pathlib import Path

from . import ncio, settings
from .functions import filter_data, interp_wrapper


def build_parser():
    parser = argparse.ArgumentParser(
        description='PGW for ERA5: preprocess GCM data (SCEN-HIST deltas and HIST climatology) before '
                    'modifying the ERA5 files: "smoothing" of daily annual cycles and/or "regridding" to '
                    'the ERA5 grid.  Input files follow ${var_name}_${file_name_base}.nc (settings.py).')
    parser.add_argument('processing_step', type=str, choices=['smoothing', 'regridding'])
    parser.add_argument('-i', '--input_dir', type=str, help='Directory with input GCM data files.')
    parser.add_argument('-o', '--output_dir', type=str, help='Directory for the preprocessed files.')
    parser.add_argument('-e', '--era5_file_path', type=str, default=None,
                        help='Example ERA5 file from which to take the target grid.')
    parser.add_argument('-v', '--var_names', type=str,
                        default='ta,hur,ua,va,zg,hurs,tas,ps,tos,ts,siconc',
                        help='Comma-separated variable names to process.')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)
    if args.input_dir is None:
        raise ValueError('Input directory (-i) is required.')
    if args.output_dir is None:
        raise ValueError('Output directory (-o) is required.')
    if (args.processing_step == 'regridding') and (args.era5_file_path is None):
        raise ValueError('era5_file_path is required for regridding step.')
    Path(args.output_dir).mkdir(exist_ok=True, parents=True)
    var_names = args.var_names.split(',')
    print('Run {} for variable names {}.'.format(args.processing_step, var_names))
    ds_era5 = ncio.open_dataset(args.era5_file_path) if args.era5_file_path else None
    for var_name in var_names:
        print(var_name)
        for clim_period in ['HIST', 'SCEN-HIST']:                  # step_02_preproc_deltas.py:116-121
            var_file_name = settings.file_name_bases[clim_period].format(var_name)
            inp_file = os.path.join(args.input_dir, var_file_name)
            out_file = os.path.join(args.output_dir, var_file_name)
            if args.processing_step == 'smoothing':
                filter_data(inp_file, var_name, out_file)
            else:
                try:
                    ds_gcm = ncio.open_dataset(inp_file)
                except Exception:
                    raise RuntimeError("Files for variable " + var_name + " are missing")
                ds_gcm = interp_wrapper(ds_gcm, ds_era5, var_name,
                                        i_use_xesmf=settings.i_use_xesmf_regridding,
                                        nan_interp_kernel_radius=settings.nan_interp_kernel_radius,
                                        nan_interp_sharpness=settings.nan_interp_sharpness)
                ds_gcm.to_netcdf(out_file)


if __name__ == "__main__":
    main()
