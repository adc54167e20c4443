"""Turn ncu output into the committed summaries under profiles/ (run in the build container, no GPU needed):
  python tools/summarize_ncu.py launches gpurun_out/launches_v2.csv  > profiles/rNN_launches_summary.csv
  python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep     > profiles/rNN_ncu_full.csv"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__cluster_size",
           "sm__cycles_elapsed.avg.per_second"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "")[:70]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v * 1e3 if unit in ("s", "second") else v
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_ms,share_pct,avg_ms")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{n},{ms:.4f},{100 * ms / tot:.2f},{ms / n:.4f}')


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(m) for m in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    print("Kernel Name," + ",".join(hdr[i] for i in idx))
    print("," + ",".join(units[i] for i in idx))
    for r in data:
        print('"' + short(r[kn]) + '",' + ",".join(r[i].replace(",", "") for i in idx))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
