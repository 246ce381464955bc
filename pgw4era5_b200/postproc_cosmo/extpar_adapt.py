"""
COSMO-specific: perturb the Extpar deep-soil temperature climatology T_CL with the annual mean of
the ts climate delta.  Drop-in for the reference's ``postproc_cosmo/extpar_adapt.py`` (same
arguments); ``load_delta(path, 'ts', None)`` + ``.mean(dim='time')`` (:20-33) run through
``functions.load_delta`` and ``pgw_time_mean_f32`` on the GPU.

    python -m pgw4era5_b200.postproc_cosmo.extpar_adapt extpar.nc -d deltas
"""
import argparse
import ctypes as C

import numpy as np
import torch

from .. import _native as N
from .. import ncio
from ..functions import load_delta

var_name_map = {
    'ts': 'T_CL',
}


def extpar_adapt(ext_file_path, delta_inp_path):
    ext_file = ncio.open_dataset(ext_file_path)
    print('update deep soil temperature')
    delta_ts = load_delta(delta_inp_path, 'ts', None)                     # full (leap-day free) series
    series = torch.as_tensor(np.ascontiguousarray(delta_ts.values, dtype=np.float32), device="cuda")
    nt = series.shape[0]
    clim = torch.empty(series.shape[1:], device="cuda", dtype=torch.float32)
    N.check(N.lib.pgw_time_mean_f32(C.c_void_p(series.data_ptr()), nt, C.c_void_p(clim.data_ptr()), clim.numel(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pgw_time_mean_f32")
    delta_ts_clim = clim.cpu().numpy()
    print(delta_ts_clim)
    t_cl = ext_file[var_name_map['ts']]
    new = np.asarray(t_cl.data) + delta_ts_clim.squeeze().reshape(np.asarray(t_cl.data).shape).astype(t_cl.data.dtype)
    ext_file[var_name_map['ts']] = ncio.Variable(t_cl.dims, new, t_cl.attrs)
    ext_file.to_netcdf(ext_file_path, mode='w')                            # the reference edits in place ('a')
    ext_file.close()
    print('Done.')


def main(argv=None):
    parser = argparse.ArgumentParser(
        description='COSMO-specific: Perturb Extpar soil temperature climatology with ts climate delta.')
    parser.add_argument('extpar_file_path', type=str, help='Path to extpar file to modify T_CL.')
    parser.add_argument('-d', '--delta_input_dir', type=str, default=None,
                        help='Directory with GCM climate deltas to be used. This directory should have a climate '
                             'delta for ts already horizontally remapped to the grid of the extpar file which can '
                             'perhaps be done with step_02_preproc_deltas.py or otherwise with CDO.')
    args = parser.parse_args(argv)
    print(args)
    extpar_adapt(args.extpar_file_path, args.delta_input_dir)


if __name__ == "__main__":
    main()
