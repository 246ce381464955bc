"""Turn an ncu report (--set full --import-source on) into the markdown summary kept in profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_column.ncu-rep > profiles/rN_column_kernel.md
"""
import collections
import csv
import json
import os
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print("# ncu summary: %s\n" % os.path.basename(rep))
    print("kernel: `%s`\n" % m.get('Kernel Name', ('?', ''))[0])
    print("| metric | value | unit |\n|---|---|---|")
    for k in WANT:
        if k in m:
            print("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
    rd = float(m['dram__bytes_read.sum'][0]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[m['dram__bytes_read.sum'][1]]
    wr = float(m['dram__bytes_write.sum'][0]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[m['dram__bytes_write.sum'][1]]
    print("\nDRAM traffic per launch: %.3f GB read + %.3f GB written = %.3f GB" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    ops, stall = collections.Counter(), collections.Counter()
    tot_i = tot_s = 0
    for r in rows[2:]:
        try:
            ni, ns = int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)
        except (ValueError, IndexError):
            continue
        s = r[ix['Source']].split()
        if not s:
            continue
        ops[(s[1] if s[0].startswith('@') else s[0]).split('.')[0]] += ni
        tot_i += ni
        tot_s += ns
        for n, i in ix.items():
            if n.startswith('stall_') and 'Not Issued' not in n:
                stall[n] += int(r[i] or 0)
    print("\nwarp instructions executed: %d" % tot_i)
    print("\nstall reasons (share of samples): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in stall.most_common(8)))
    print("\nopcode mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in ops.most_common(20)))
    # bench.py reads the DRAM traffic of the COLUMN kernel from this file (roofline.traffic): only a capture of that
    # kernel may update it
    if "pgw_column" in m.get('Kernel Name', ('', ''))[0]:
        json.dump({"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": os.path.basename(rep)},
                  open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "column_kernel_traffic.json"), "w"))


if __name__ == "__main__":
    main(sys.argv[1])
