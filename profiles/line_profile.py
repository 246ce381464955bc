"""Per-source-line dynamic instruction counts / stall samples of a kernel.

Joins the SASS page of an ncu report (--set full --import-source on) with nvdisasm -g
line info of the cubin the report was taken from (same instruction order).

    python profiles/line_profile.py gpurun_out/prof.ncu-rep pgw4era5_b200/csrc/pgw_timestep.o 'pgw_column_kernelILi128ELb1'
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, func_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur, active = [], None, False
    for line in txt.splitlines():
        if line.startswith('.text.'):
            active = func_pat in line
            continue
        if not active:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            out.append((cur, m.group(2)))
    return out


def main(rep, obj, func_pat, src=None):
    lines = sass_lines(obj, func_pat)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    body = [r for r in rows[2:] if len(r) > ix['Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
    if len(body) != len(lines):
        sys.stderr.write("warning: %d profiled vs %d disassembled instructions\n" % (len(body), len(lines)))
    inst, samp = collections.Counter(), collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    scols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    for (loc, _), r in zip(lines, body):
        inst[loc] += int(r[ix['Instructions Executed']])
        samp[loc] += int(r[ix['# Samples']] or 0)
        for n in scols:
            stall[loc][n[6:]] += int(r[ix[n]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    text = {}
    if src:
        text = {i + 1: l.rstrip() for i, l in enumerate(open(src))}
    print("total warp instructions %d, samples %d" % (ti, ts))
    for loc in sorted(inst, key=lambda x: (x is None, x)):
        if inst[loc] < ti * 0.001 and samp[loc] < ts * 0.001:
            continue
        f, l = loc if loc else ('?', 0)
        t = text.get(l, '')[:90] if f and src and f in src else f
        top = ' '.join('%s:%d' % (k, 100 * v // max(samp[loc], 1)) for k, v in stall[loc].most_common(2))
        print("%4d %6.2f%% inst %6.2f%% smp  %-22s %s" % (l, 100.0 * inst[loc] / ti, 100.0 * samp[loc] / max(ts, 1), top, t))


if __name__ == '__main__':
    main(*sys.argv[1:])
