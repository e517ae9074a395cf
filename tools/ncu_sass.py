"""SASS of one kernel in address order with per-instruction execution counts, from
`ncu -i rep --page source --csv --print-source cuda,sass`:   python tools/ncu_sass.py src.csv > listing.txt"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    cur_file, idx, line_no = None, None, None
    out = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            idx = {name: i for i, name in reversed(list(enumerate(r)))}
            continue
        if idx is None:
            continue
        if r[0] != "":
            line_no = r[0]
            continue
        try:
            addr = int(r[2], 16)
            inst = int(r[idx["Instructions Executed"]])
            thr = int(r[idx["Thread Instructions Executed"]])
            smp = int(r[idx["# Samples"]])
        except (ValueError, KeyError, IndexError):
            continue
        out[addr] = (inst, thr, smp, r[3].strip(), f"{cur_file}:{line_no}")
    base = min(out)
    tot = sum(v[0] for v in out.values())
    print(f"# {len(out)} instructions, {tot} warp-level executions")
    for addr in sorted(out):
        inst, thr, smp, sass, where = out[addr]
        print(f"{addr - base:6x} {inst:10d} {thr / max(inst, 1):5.1f} {smp:6d}  {sass:<70s} {where}")


if __name__ == "__main__":
    main(sys.argv[1])
