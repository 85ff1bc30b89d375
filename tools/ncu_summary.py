"""Turn ncu output brought back from the GPU box into the markdown kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv        # per-kernel shares of the launch list
    python tools/ncu_summary.py full gpurun_out/x_prof.ncu-rep [...]      # key metrics of a --set full capture

Only reads files; ncu itself is run on the GPU box (see profiles/*.md for the exact commands).
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg.per_second",
]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:70]


def launches(path):
    with open(path) as fh:
        text = "".join(l for l in fh if l.startswith('"'))
    agg = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO(text)):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        k = short(r["Kernel Name"])
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.001:
            continue
        print(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |")


def full(path):
    res = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(res.stdout)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"\n`{short(r[hdr.index('Kernel Name')])}` ({path})\n\n| metric | value |\n|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} [{units[i]}] | {r[i]} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[2:]:
            full(p)
