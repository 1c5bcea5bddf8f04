#!/usr/bin/env python
"""Attribute ncu's per-instruction samples of one kernel to source lines and (inlined) functions.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
  cuobjdump -xelf all libsim5b200.so ; nvdisasm -gi -c capi.sm_100a.cubin > all_gi.txt
  python tools/ncu_hotspots.py src.csv all_gi.txt <mangled-kernel-substring> [--top 30]

For every SASS instruction the nvdisasm listing gives the innermost source line and its inlined-at chain; the ncu
CSV gives samples / executed counts in the same instruction order.  Output: samples and executed warp instructions
by top-level pixel-pipeline line (the outermost frame in pixel.cuh / kernels.cuh) and by innermost function file:line.
"""
import csv
import re
import sys
from collections import defaultdict


def parse_disasm(path, kernel):
    ins = []
    cur = None
    active = False
    pending = []
    for line in open(path, errors="ignore"):
        if line.startswith(".text."):
            active = kernel in line
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
            for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
                chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
            if pending and pending_is_inline[0]:
                pending.extend(chain)
            else:
                pending[:] = chain
            pending_is_inline[0] = "inlined at" in line
            cur = list(pending)
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
            pending_is_inline[0] = False
            pending.clear() if False else None
    return ins


pending_is_inline = [False]


def main():
    src_csv, disasm, kernel = sys.argv[1:4]
    top = 30
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    ins = parse_disasm(disasm, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    print("instructions: disasm %d, ncu %d" % (len(ins), len(data)))
    n = min(len(ins), len(data))
    by_inner = defaultdict(lambda: [0, 0])
    by_outer = defaultdict(lambda: [0, 0])
    by_file = defaultdict(lambda: [0, 0])
    tot_s = tot_e = 0
    mism = 0
    for k in range(n):
        addr, text, chain = ins[k]
        r = data[k]
        if text.split()[0].lstrip("@!P0123456789 ") not in r[col["Source"]] and text.split()[-1] not in r[col["Source"]]:
            mism += 1
        s = int(r[col["# Samples"]] or 0)
        e = int(r[col["Instructions Executed"]] or 0)
        tot_s += s
        tot_e += e
        if not chain:
            chain = [("?", 0)]
        inner = chain[0]
        outer = next((c for c in reversed(chain) if c[0] in ("pixel.cuh",)), chain[-1])
        by_inner[inner][0] += s; by_inner[inner][1] += e
        by_outer[outer][0] += s; by_outer[outer][1] += e
        by_file[inner[0]][0] += s; by_file[inner[0]][1] += e
    print("text mismatches:", mism, " total samples", tot_s, " warp instructions", tot_e)
    for title, d in (("by file of the innermost frame", by_file), ("by outermost pixel.cuh line", by_outer), ("by innermost line", by_inner)):
        print("\n== %s ==" % title)
        for key, (s, e) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print("%6.2f%% samples  %6.2f%% instr   %s" % (100.0 * s / max(tot_s, 1), 100.0 * e / max(tot_e, 1), key))


if __name__ == "__main__":
    main()
