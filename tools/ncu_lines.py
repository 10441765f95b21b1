#!/usr/bin/env python
"""Development helper: attribute ncu warp-stall samples to source lines.

  ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g search_kernel.sm_100a.cubin > lines.txt
  python tools/ncu_lines.py sass.csv lines.txt '<mangled kernel name>' [source.cu]
"""
import collections
import csv
import re
import sys

sass_csv, lines_txt, kernel = sys.argv[1:4]
src = open(sys.argv[4]).read().split("\n") if len(sys.argv) > 4 else None

# address -> line from nvdisasm -g
addr_line = {}
cur_line = None
inside = False
for ln in open(lines_txt):
    if ln.startswith(".text."):
        inside = ln.strip() == f".text.{kernel}:"
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = "inlined" in ln
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur_line:
        addr_line[int(m.group(1), 16)] = cur_line

rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ai, si, ni = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
ii = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_line = collections.Counter()
inst_line = collections.Counter()
stall_line = collections.defaultdict(collections.Counter)
base = None
total = 0
for r in rows[2:]:
    if len(r) <= ni:
        continue
    try:
        a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai])
    except ValueError:
        continue
    if base is None:
        base = a
    off = a - base
    n = int(r[ni] or 0)
    total += n
    key = addr_line.get(off, ("?", 0))
    by_line[key] += n
    inst_line[key] += int(r[ii] or 0)
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            stall_line[key][h[6:]] += v
print(f"total samples {total}, mapped lines {len(by_line)}")
for (f, l), n in by_line.most_common(45):
    text = src[l - 1].strip()[:90] if src and f.endswith("search_kernel.cu") and 0 < l <= len(src) else ""
    top = ", ".join(f"{k}:{v}" for k, v in stall_line[(f, l)].most_common(3))
    print(f"{100*n/total:5.1f}%  {f}:{l:<4} inst={inst_line[(f,l)]:<9} [{top}]  {text}")
