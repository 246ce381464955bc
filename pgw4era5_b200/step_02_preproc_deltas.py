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
from .parallel import IterMP


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
    # not in the reference (which processes the files one after the other on the CPU): the
    # (variable, HIST | SCEN-HIST) files are independent, so they are dealt out to worker processes,
    # worker i on GPU i % device_count
    parser.add_argument('-p', '--n_par', type=int, default=1,
                        help='Number of worker processes (one per GPU) over which the files are distributed.')
    return parser


def process_file(processing_step, inp_file, out_file, var_name, era5_file_path):
    """One (variable, climate period) file: step_02_preproc_deltas.py:128-149."""
    if processing_step == 'smoothing':
        filter_data(inp_file, var_name, out_file)
        return out_file
    ds_era5 = ncio.open_dataset(era5_file_path)
    try:
        ds_gcm = ncio.open_dataset(inp_file)
    except Exception:
        raise RuntimeError("Files for variable " + var_name + " are missing")
    ds_gcm = interp_wrapper(ds_gcm, ds_era5, var_name,
                            i_use_xesmf=settings.i_use_xesmf_regridding,
                            nan_interp_kernel_radius=settings.nan_interp_kernel_radius,
                            nan_interp_sharpness=settings.nan_interp_sharpness)
    ds_gcm.to_netcdf(out_file)
    return out_file


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
    tasks = []
    for var_name in var_names:
        print(var_name)
        for clim_period in ['HIST', 'SCEN-HIST']:                  # step_02_preproc_deltas.py:116-121
            var_file_name = settings.file_name_bases[clim_period].format(var_name)
            tasks.append(dict(inp_file=os.path.join(args.input_dir, var_file_name),
                              out_file=os.path.join(args.output_dir, var_file_name), var_name=var_name))
    IMP = IterMP(njobs=args.n_par, run_async=True)
    IMP.run(process_file, dict(processing_step=args.processing_step, era5_file_path=args.era5_file_path), tasks)
    return IMP.output


if __name__ == "__main__":
    main()
