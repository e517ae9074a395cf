"""Turns ncu output into the small tracked summaries under profiles/ (run here, on reports brought back in gpurun_out/).

    python tools/ncu_summary.py kernel REP.ncu-rep OUT.json [note]      # key metrics + stall reasons of every kernel in the report
    python tools/ncu_summary.py traffic REP.ncu-rep OUT.json WORKLOAD   # DRAM bytes per launch (bench.py's roofline.traffic)
    python tools/ncu_summary.py launches LIST.csv OUT_summary.csv       # per-kernel totals and shares of a launch list
    python tools/ncu_summary.py hotlines REP.ncu-rep OUT.txt [top]      # per-source-line instruction / sample shares + phases
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernel(rep, out_path, note=""):
    hdr, units, rows = raw_rows(rep)
    result = {}
    for r in rows:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        d = {"_report": rep.split("/")[-1]}
        if note:
            d["_note"] = note
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
        d["warps_stalled_per_issue_active"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        result[name] = d
    json.dump(result, open(out_path, "w"), indent=1)
    print(out_path, list(result))


def traffic(rep, out_path, workload):
    hdr, units, rows = raw_rows(rep)
    r = rows[0]

    def b(key):
        i = hdr.index(key)
        return int(float(r[i].replace(",", "")) * UNIT_SCALE[units[i]])

    rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
    json.dump({"kernel": r[hdr.index("Kernel Name")].split("(")[0], "workload": workload, "dram_bytes_read": rd,
               "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "source": f"ncu --set full --clock-control none, one launch ({rep.split('/')[-1]}; summary in profiles/)"},
              open(out_path, "w"), indent=1)
    print(out_path, rd + wr)


def launches(list_csv, out_csv):
    rows = list(csv.reader(l for l in open(list_csv) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out_csv, "w") as f:
        f.write("kernel,launches,total_us,avg_us,share_of_listed_time\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{c},{t / 1e3:.1f},{t / c / 1e3:.2f},{t / total:.4f}\n")
    print(out_csv, len(agg), "kernels")


def hotlines(rep, out_path, top=45):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, idx = None, None, None
    agg = collections.defaultdict(lambda: [0, 0, 0, "", collections.Counter()])
    stall_names = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Line No":
            hdr = r
            idx = {n: i for i, n in reversed(list(enumerate(hdr)))}
            stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            inst, thr, smp = int(r[idx["Instructions Executed"]]), int(r[idx["Thread Instructions Executed"]]), int(r[idx["# Samples"]])
        except ValueError:
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += inst
        a[1] += thr
        a[2] += smp
        a[3] = r[1].strip()[:96]
        for sname in stall_names:
            try:
                a[4][sname] += int(r[idx[sname]])
            except ValueError:
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    stalls = collections.Counter()
    for a in agg.values():
        stalls.update(a[4])
    with open(out_path, "w") as f:
        f.write(f"{rep.split('/')[-1]}: {tot} warp instructions, {tots} samples\n")
        f.write("stall samples: " + ", ".join(f"{k[6:]} {v / max(1, sum(stalls.values())) * 100:.1f}%" for k, v in stalls.most_common(10)) + "\n")
        per_file = collections.Counter()
        for (fn, _), a in agg.items():
            per_file[fn] += a[0]
        f.write("instructions by file: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in per_file.most_common(8)) + "\n\n")
        for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            top_stall = ", ".join(f"{k[6:]} {v}" for k, v in a[4].most_common(2))
            f.write(f"{a[0] / tot * 100:5.1f}% inst {a[2] / tots * 100:5.1f}% smp  thr/inst {a[1] / max(1, a[0]):4.1f}  {fn}:{ln:<4} {a[3]}   [{top_stall}]\n")
        f.write("\nby samples:\n")
        for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:20]:
            top_stall = ", ".join(f"{k[6:]} {v}" for k, v in a[4].most_common(2))
            f.write(f"{a[2] / tots * 100:5.1f}% smp {a[0] / tot * 100:5.1f}% inst  {fn}:{ln:<4} {a[3]}   [{top_stall}]\n")
    print(out_path)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "kernel":
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif mode == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    elif mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif mode == "hotlines":
        hotlines(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 45)
