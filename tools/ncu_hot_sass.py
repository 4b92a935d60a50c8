#!/usr/bin/env python
"""Rank the SASS instructions of one kernel by warp-stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:X --launch-count 1 > src.csv
       ncu_hot_sass.py src.csv [topN]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    iS, iSrc, iAddr, iEx = (hdr.index(k) for k in ("# Samples", "Source", "Address", "Instructions Executed"))
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iS]) for r in data)
    base = int(data[0][iAddr], 16)
    print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", "| samples", tot, "| instructions", len(data))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:top_n]:
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{int(r[iAddr], 16) - base:6x} {int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% ex={r[iEx]:>9s} "
              f"{r[iSrc].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main()
