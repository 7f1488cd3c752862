#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: per source line (with inlining chain),
share of executed warp instructions and of stall samples.

  ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
  cuobjdump -xelf all libtfcfft.so; nvdisasm -g tfcfft_api.sm_100a.cubin > all.sass
  python tools/ncu_lines.py src.csv all.sass <mangled kernel name> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    """[(line_key, sass_text)] for the kernel's instructions in order."""
    out, on, cur = [], False, ("?", 0)
    stack = ""
    for ln in open(path, errors="replace"):
        if ln.startswith(".text." + kernel + ":"):
            on = True
            continue
        if on and ln.startswith("//-----"):
            break
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = m.group(3)
            cur = (m.group(1).split("/")[-1], int(m.group(2)), inl.strip())
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    src, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kernel_match(r[1], kernel))
    hdr = rows[start + 1]
    ie, ss, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    data = []
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        data.append(r)
    lines = sass_lines(sass, kernel)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} disassembled vs {len(data)} profiled instructions", file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    ti = tsamp = 0.0
    for (key, _), r in zip(lines, data):
        i, s = float(r[ie] or 0), float(r[ss] or 0)
        a = agg[key[:2]]
        a[0] += i
        a[1] += s
        a[2] += 1
        ti += i
        tsamp += s
    print(f"total warp instructions {ti:.0f}, stall samples {tsamp:.0f}, sass instructions {len(data)}")
    print(" inst%  stall%  nsass  file:line")
    for key, (i, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*i/ti:6.2f} {100*s/tsamp:7.2f} {n:6d}  {key[0]}:{key[1]}")


def kernel_match(pretty, mangled):
    name = re.sub(r"^_ZN6tfcfft\d+", "", mangled)
    name = re.match(r"[a-z_0-9]+", name).group(0)
    return name in pretty


if __name__ == "__main__":
    main()
