"""Aggregates `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > src.csv
    python tools/ncu_hotlines.py src.csv [top_n]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])
    line_no, line_src = None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            idx = {name: i for i, name in reversed(list(enumerate(hdr)))}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":
            line_no, line_src = r[0], r[1]
        try:
            inst = int(r[idx["Instructions Executed"]])
            thr = int(r[idx["Thread Instructions Executed"]])
            smp = int(r[idx["# Samples"]])
        except (ValueError, KeyError):
            continue
        key = (cur_file, int(line_no))
        a = agg[key]
        a[0] += inst
        a[1] += thr
        a[2] += smp
        a[3] = line_src
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0] / tot * 100:5.1f}% inst {a[2] / tots * 100:5.1f}% smp  thr/inst {a[1] / max(1, a[0]):4.1f}  {f}:{ln:<4} {a[3].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
