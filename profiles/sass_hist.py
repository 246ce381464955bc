#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of an object file / shared library:

    python profiles/sass_hist.py pgw4era5_b200/csrc/pgw_column_tma.o 'Lb1ELi137ELi56' [--top 40]

Used for the evidence that the TMA flavour really is a TMA/mbarrier kernel (UTMALDG, UTMASTG, SYNCS) and to
compare instruction counts of kernel variants before spending GPU time."""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, funcs = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        h = collections.Counter()
        for _, text in ins:
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            h[op.split(".")[0]] += 1
        print("%s: %d instructions" % (name, len(ins)))
        # backward branches = loops; print their spans
        for addr, text in ins:
            m = re.search(r"BRA.*?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                tgt = int(m.group(1), 16)
                n = sum(1 for a, _ in ins if tgt <= a <= addr)
                print("  loop 0x%x..0x%x: %d instructions" % (tgt, addr, n))
        for op, n in h.most_common(top):
            print("  %-12s %5d" % (op, n))


if __name__ == "__main__":
    main()
