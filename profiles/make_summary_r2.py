"""Regenerate profiles/r2_summary.md from the JSON files of round 2 in profiles/."""
import json
import os

P = os.path.dirname(os.path.abspath(__file__))
L = lambda f: json.load(open(os.path.join(P, f)))
m = L('r2_multigpu.json')
b1, ref, b2, b8 = m['n1']['bench'], m['n1']['reference_arm'], m['n2']['bench'], m['n8']['bench']
p8, p2 = m['n8']['bench_with_peer_memory_band_exchange'], m['n2']['bench_with_peer_memory_band_exchange']
b4 = m['n4']['bench']
s2, f, sf, rd, pg = L('r2_step02.json'), L('r2_files.json'), L('r2_step02_files.json'), L('r2_ref_dtypes.json'), L('r2_parity_global.json')
mx = lambda k: max(c['maxerr'][k] for c in pg['cases'])
sb = m['n8']['step02_banded']
eu, gl = f['european_domain'], f['global']
txt = f'''# Round 2 - measured numbers (B200, sm_100a, SM clock {b1['clocks']['sm_mhz']:.0f} MHz, throttle reasons: {b1['clocks']['reasons']})

All numbers from `bench.py` / `tests/bench_*.py` / `tests/multigpu_*.py` on `gpurun` boxes of this pool (1, 2 and 8 B200);
the JSON lines they come from are in `r2_multigpu.json`, `r2_step02.json`, `r2_step02_files.json`, `r2_files.json`,
`r2_parity_global.json`, `r2_ref_dtypes.json`; ncu evidence: `r2_column_kernel.md`, `r2_regrid_walk_kernel.md`,
`r2_regrid_rows_kernel.md` (the round-1 kernel at full size), `r2_smooth_kernel.md`, `r2_launches.csv`, `r2_sass_tma.txt`.
Workload: BASELINE configs[1], global 0.25 degree ERA5 timestep (721x1440 columns x 137 levels, plev19 monthly deltas),
inputs cycling through 4 distinct device-resident timesteps (2.3 GB each >> 126 MB L2), driver schedule (20 steps, 5 warm-up).

| quantity | round 1 | round 2 | source |
|---|---|---|---|
| timesteps/s, 1xB200 (`value`) | 760 | **{b1['value']:.1f}** | `python bench.py --steps 20 --warmup 5` |
| ms per timestep (whole step) | 1.316 | {b1['ms_per_step']:.4f} | same |
| column kernel per launch (events in the timed region; union of the intervals / launches) | 1.285 | {b1['roofline']['kernel_ms']:.4f} ms (mean of the individual launch durations with two streams: {b1['roofline']['kernel_ms_mean_of_launch_durations']:.3f}) | same |
| achieved / measured HBM peak 6 535 GB/s (`roofline.frac`) | 0.632 | **{b1['roofline']['frac']:.3f}** ({b1['roofline']['achieved']:.0f} GB/s); whole step charged: {b1['roofline']['frac_whole_step']:.3f} | same |
| DRAM traffic per launch (ncu) | 5.305 GB | {b1['roofline']['traffic']/1e9:.3f} GB = {b1['roofline']['traffic']/b1['roofline']['algorithmic_bytes']:.3f} x algorithmic | `r2_column_kernel.md` |
| warp instructions per launch / issue slots active | 839 M / 57 % | 796 M / 56 % | `r2_column_kernel.md` |
| host time per submit; `ms_per_step - kernel_ms` | - ; 0.031 ms | {b1['config']['host_us_per_submit']:.0f} us; {b1['ms_per_step']-b1['roofline']['kernel_ms']:.3f} ms | bench line |
| `e2e` (host buffers, copies inside the timing), 1 GPU | 19.8 | {b1['e2e']['value']:.1f} timesteps/s = {b1['e2e']['frac_of_link']:.2f} x the measured link bound ({b1['e2e']['link_bound']['gb_per_s_per_direction_per_gpu']:.1f} GB/s per direction with both directions busy) | same |
| timesteps/s, 2xB200 / 4xB200 | 1 524 / 3 042 | {b2['value']:.0f} / {b4['value']:.0f} | torchrun, 20 steps |
| timesteps/s, 8xB200; efficiency vs 8 x N=1 | 5 562; 0.91 | **{p8['value']:.0f}**; {p8['value']/8/b1['value']:.3f} | torchrun, 20 steps |
| `ms_per_step - kernel_ms` at N = 8 | 0.157 ms | {b8['ms_per_step']-b8['roofline']['kernel_ms']:.3f} ms | same |
| NCCL broadcast of the climatology (5.03 GB), 2 / 8 GPUs | 355 / 1 068 ms (8 collectives, incl. NCCL start-up) | {b2['config']['broadcast']['ms']:.1f} ms ({b2['config']['broadcast']['gb_per_s']:.0f} GB/s) / {b8['config']['broadcast']['ms']:.1f} ms ({b8['config']['broadcast']['gb_per_s']:.0f} GB/s), one collective | `config.broadcast` |
| `e2e` at N = 2 / 4 / 8 (sum over ranks) | 22.8 / 22.0 / 28.2 | {b2['e2e']['value']:.1f} / {b4['e2e']['value']:.1f} / {b8['e2e']['value']:.1f} = {b2['e2e']['frac_of_link']:.2f} / {b4['e2e']['frac_of_link']:.2f} / {b8['e2e']['frac_of_link']:.2f} x the platform bound measured bare ({b2['e2e']['link_bound']['gb_per_s_per_direction_per_gpu']:.1f} / {b4['e2e']['link_bound']['gb_per_s_per_direction_per_gpu']:.1f} / {b8['e2e']['link_bound']['gb_per_s_per_direction_per_gpu']:.1f} GB/s per direction and GPU when all ranks copy both ways at once) | `e2e.link_bound`, `tests/multigpu_pcie.py` |
| one global snapshot in latitude bands, plev37, thresh 1e-3 (BASELINE configs[4]), 8 / 2 GPUs | 0.65 / 1.15 ms | **{p8['latband']['ms_per_snapshot']:.3f}** / {p2['latband']['ms_per_snapshot']:.3f} ms with the exchange fused over peer memory ({p8['latband']['nccl_form_of_the_exchange']['ms_per_snapshot']:.3f} / {p2['latband']['nccl_form_of_the_exchange']['ms_per_snapshot']:.3f} with one NCCL all-reduce); band kernels alone {p8['latband']['ms_per_snapshot_band_kernels_only']:.3f} ms; 4 GPUs: {b4['config']['latband']['ms_per_snapshot']:.3f} ms; {p8['latband']['n_iter']} iterations = whole grid; band vs whole-grid fields max abs diff {p8['latband']['band_vs_whole_grid_max_abs_diff']} | `config.latband` |
| CPU arm: oracle port on 16 host cores, IterMP-driven | 0.105 (17.75 % sample) | {ref['value']:.4f} timesteps/s ({ref['config']['sampled_fraction_of_timestep']:.3f} of a timestep per step) | `bench.py --impl reference` |
| step_02 regridding, one 3-D daily variable (28.8 GB out) | 8.75 ms (0.54) | **{s2['regridding']['ms']:.2f} ms = {s2['regridding']['achieved_gbs']:.0f} GB/s ({s2['regridding']['frac_of_peak']:.3f} of peak)** | `tests/bench_step02.py` |
| step_02 smoothing, one 3-D daily variable | 1.05 ms (0.53) | {s2['smoothing']['ms']:.3f} ms ({s2['smoothing']['frac_of_peak']:.3f}) | same |
| step_02 regridding on 8 GPUs by target latitude | - | broadcast {sb['broadcast_source_ms']:.1f} + smoothing {sb['smoothing_ms_replicated']:.1f} + band kernels {sb['regrid_band_ms']:.1f} ms; gathering a tenth of the result {sb['broadcast_plus_gather_regrid_of_684_fields_ms']:.1f} ms | `tests/multigpu_step02.py` |
| step_02 file -> file (48 days, 2 files) | - | regridding {sf['regridding_cli_s_for_2_files']:.1f} s, of which CUDA kernels {100*sf['kernel_share_of_regridding_wall']:.3f} %; {sf['regridding_output_gb_per_s']:.2f} GB/s of NetCDF output | `tests/bench_step02_files.py` |
| step_03 file -> file, EU files (126 MB) | 23.9 files/s | {eu['pipelined_files_per_s']:.1f} files/s; stages per file: read {1e3*eu['stages']['read_file_to_pinned_s']:.1f} ms, H2D+pass+D2H {1e3*eu['stages']['h2d_kernel_d2h_s']:.1f} ms, write {1e3*eu['stages']['write_pinned_to_file_s']:.1f} ms | `tests/bench_files.py --breakdown` |
| step_03 file -> file, global files (2.3 GB), 3 files | 1.05 files/s | {gl['pipelined_files_per_s']:.2f} files/s; stages per file: read {gl['stages']['read_file_to_pinned_s']:.3f} s, H2D+pass+D2H {gl['stages']['h2d_kernel_d2h_s']:.3f} s, **write {gl['stages']['write_pinned_to_file_s']:.3f} s** (the file system of the box: {gl['stages']['file_bytes']/gl['stages']['write_pinned_to_file_s']/1e9:.1f} GB/s); the slowest stage allows {gl['stages']['slowest_stage_files_per_s']:.2f} files/s | same |
| parity on ALL 1 038 240 columns vs the oracle, configs[1], [2] (3 dates), [4] | 4 rows, T/U/V only | ps <= {mx('PS'):.1e} Pa, T <= {mx('T'):.1e} K, QV <= {mx('QV'):.1e}; iteration counts {', '.join(str(c['n_iter_gpu'])+'/'+str(c['n_iter_oracle']) for c in pg['cases'])}; 0 columns beyond tolerance | `r2_parity_global.json` |
| reference-dtype mode vs oracle(emulate_file_dtypes), {rd['runs']} runs, {rd['columns_compared']} columns | - | iteration counts identical in all runs; ps bit-identical on {100*rd['ps_bit_identical_fraction']:.2f} %, within 1e-2 Pa on {100*rd['ps_within_1e-2_Pa_fraction']:.2f} %, max {rd['ps_max_abs_diff_Pa']:.4f} Pa | `r2_ref_dtypes.json` |

Launch list of one bench run (`r2_launches.csv`, ncu `--metrics gpu__time_duration.sum --clock-control none -k regex:pgw_|time_mean`;
cold-cache, serialised, 13 timesteps): `pgw_column_tma_kernel<1, 137, 56>` 1 263 us x13 (95.2 %); `pgw_rewrite_kernel` 58 us x13 (4.4 %:
four warm-up launches that over-predicted `k_spec` really rewrite, ~170 us each, the others return at once);
`pgw_converge_kernel` 3 us (0.2 %); `pgw_timestep_init_kernel` 2 us (0.2 %).  The column kernel's share of the step agrees
with the event timing (kernel 1.215 of 1.223 ms).

Column kernel this round (ms per launch, global/plev19): 1.295 (round-1 kernel on this round's boxes) -> 1.251-1.271
(packed float32 pairs: FADD2/FMUL2/FFMA2 for the arithmetic the two levels of a pair share); per step 1.314 -> 1.269 (one C
call per timestep, cached tensor maps and argument block, four timesteps in flight) -> 1.222 (consecutive timesteps on two
streams: the partly filled last wave of a launch overlaps the next launch).  Measured and dropped: unrolling the streamed /
parked / both sweep loops on top of the packing (1.32 / 1.35 / 1.46 ms), five ring slots (1.61 ms), three streams (no gain
over two).  Regrid kernel: 8.70 (round-1 rows kernel) -> 7.00 (walking kernel: staged source rows, row table, uniform
three-column pattern) -> 6.51 (4 rows per barrier, cp.async staging, 16-row chunks) -> 6.22 (32-row chunks); dropped: float64
staging (6.46), float64 staging + software-pipelined passes (7.88), 2 CTAs/SM (8.2).  Smoothing: 1.04 -> 1.006 (8 loads in
flight per thread) -> 0.953 (cp.async ring, 16 stamps ahead); two points per thread 1.31, constant-memory tables 1.50, other
launch shapes 0.98-1.05; read pass alone 0.58 ms, write pass alone 0.42 ms; same GB/s for 8 plevs (2 MB rows) as for 19.

Tools added: `sass_hist.py` (static opcode histogram / loop sizes of a kernel), `make_summary_r2.py` (this file),
`gpu_r2*.sh` (the GPU calls of this round).
'''
open(os.path.join(P, 'r2_summary.md'), 'w').write(txt)
