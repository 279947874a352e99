"""Turn ncu outputs into the small text summaries kept under profiles/.
  launches <launches.csv>          per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
  full <report.ncu-rep>            headline metrics per captured launch of an `ncu --set full` report"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def short(name):
    m = re.search(r"(?:anonymous namespace|unnamed)\W+(k_[a-z_]+)<([^>]*)>", name)
    if m:
        return f"paa::{m.group(1)}<{m.group(2).replace('(int)', '').replace('(bool)', '')}>"
    return re.sub(r"<.*", "", name).replace("void ", "")[:70]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, total, order = collections.OrderedDict(), 0.0, []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        k = short(row["Kernel Name"])
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
        total += t
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {total / 1e3:.3f} ms summed device time "
          "(cold-cache, serialised by ncu: compare shares)")
    print(f"{'us':>12s} {'share':>7s} {'n':>5s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total > 0.002 or k.startswith("paa::"):
            print(f"{t:12.1f} {100 * t / total:6.2f}% {n:5d}  {k}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", short(r[idx["Kernel Name"]]), " grid", r[idx.get("Grid Size", 0)], " block", r[idx.get("Block Size", 0)])
        for k in KEYS:
            if k in idx:
                print(f"  {k:72s} {r[idx[k]]:>18s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
