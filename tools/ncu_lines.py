"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > dump.csv
       python tools/ncu_lines.py dump.csv [top]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fname, hdr, items = "", None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if "Instructions Executed" in r:
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            inst = int(r[hdr.index("Instructions Executed")])
            samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        items.append((inst, samp, fname, r[0], r[1].strip()[:100]))
    ti = sum(i[0] for i in items) or 1
    ts = sum(i[1] for i in items) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    print("by instructions:")
    for inst, samp, f, ln, src in sorted(items, reverse=True)[:top]:
        print(f"{inst / ti * 100:5.1f}% inst {samp / ts * 100:5.1f}% smp  {f}:{ln}: {src}")
    print("by stall samples:")
    for inst, samp, f, ln, src in sorted(items, key=lambda t: -t[1])[:top]:
        print(f"{inst / ti * 100:5.1f}% inst {samp / ts * 100:5.1f}% smp  {f}:{ln}: {src}")


if __name__ == "__main__":
    main()
