"""profiles/<round>_ncu_summary.md (+ copies of the raw evidence) from what tools/collect_evidence.sh brought back.

    python tools/make_profile_md.py r1        # reads gpurun_out/ev_*, writes profiles/r1_*
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary as NS  # noqa: E402


def capture(fn, *a):
    buf, old = io.StringIO(), sys.stdout
    sys.stdout = buf
    try:
        fn(*a)
    finally:
        sys.stdout = old
    return buf.getvalue()


def raw_rows(path):
    res = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(res.stdout)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, rows[1]))


def fnum(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def gb(d, units, k):
    v = fnum(d, k)
    u = units.get(k, "")
    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1.0)


def ms(d, units, k="gpu__time_duration.sum"):
    v = fnum(d, k)
    return v * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(units.get(k, ""), 1.0)


def train_step_table(path):
    with open(path) as fh:
        text = "".join(l for l in fh if l.startswith('"'))
    rows = [r for r in csv.DictReader(io.StringIO(text)) if r["Metric Name"] == "gpu__time_duration.sum"]
    idx = [i for i, r in enumerate(rows) if "adam" in r["Kernel Name"]]
    a, b = idx[-2] + 1, idx[-1] + 1
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        k = NS.short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] in ("ns", "nsecond") else 1)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    out = [f"One training step = {b - a} launches, {tot / 1e3:.2f} ms summed kernel time (serialised, cold cache).\n",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot >= 0.004:
            out.append(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |")
    return "\n".join(out)


def main(tag):
    os.makedirs(P, exist_ok=True)
    for src, dst in (("ev_launches.csv", f"{tag}_launches_bench_batch128.csv"),
                     ("ev_train_launches.csv", f"{tag}_launches_train_step_batch256.csv"),
                     ("ev_breakdown.txt", f"{tag}_step_breakdown_batch1024_cuda_events.txt"),
                     ("ev_configs.json", f"{tag}_secondary_configs_1gpu.json"),
                     ("ev_probe.log", f"{tag}_kernel_probes.txt")):
        if os.path.exists(os.path.join(O, src)):
            shutil.copy(os.path.join(O, src), os.path.join(P, dst))
    with open(os.path.join(O, "ev_bench.log")) as fh:
        line = [l for l in fh if l.startswith("{")][-1]
    with open(os.path.join(P, f"{tag}_bench_batch1024.json"), "w") as fh:
        fh.write(line)
    bench = json.loads(line)
    md = [f"# {tag} — ncu evidence (B200, sm_100a)\n",
          "Every ncu pass re-ran a command that had just exited 0 without ncu (`tools/collect_evidence.sh`). Numbers printed "
          "under ncu are never bench values; the bench line is `" + f"{tag}_bench_batch1024.json` (plain run, CUDA events): "
          f"**{bench['value']:.2f} samples/s**, {bench['ms_per_step']:.2f} ms per step, e2e {bench['e2e']['value']:.2f} "
          f"samples/s, SM clock {bench['clocks']['sm_mhz']} MHz, throttle reasons {bench['clocks']['reasons']}.\n",
          "## 1. Launch list of the sampling step\n",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 3 "
          "--batch 128 --no-cpu-baseline --no-e2e` (raw: `" + f"{tag}_launches_bench_batch128.csv`). Cold-cache, serialised "
          "per-launch times: compare SHARES with the CUDA-event breakdown, not absolutes. `at::` kernels are the one-off "
          "weight repack before the loop, not the step.\n",
          capture(NS.launches, os.path.join(O, "ev_launches.csv"))]
    with open(os.path.join(O, "ev_breakdown.txt")) as fh:
        bd = [l.split(" ms ", 1) for l in fh.read().strip().splitlines()]
    tot = float(bd[-1][0])
    share = collections.OrderedDict()
    for t, name in bd[:-1]:
        key = name.strip().split(" ")[0]
        share[key] = share.get(key, 0.0) + float(t)
    md.append(f"CUDA-event breakdown of one eager step at the full bench size (2048 images per forward, "
              f"`{tag}_step_breakdown_batch1024_cuda_events.txt`), total {tot:.2f} ms: " +
              ", ".join(f"{k} {v:.2f} ms ({100 * v / tot:.1f} %)" for k, v in sorted(share.items(), key=lambda kv: -kv[1])[:6]) +
              ".\n")
    md.append("## 2. `ncu --set full --clock-control none --import-source on` captures at the FULL bench size\n")
    rows, units = raw_rows(os.path.join(O, "ev_prof_conv_sw.ncu-rep"))
    md.append("`-k regex:conv3x3_sw -s 36 -c 5` on `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e` "
              "(launches 37-41 of the process; 2048 images each):\n")
    md.append("| launch | time ms | DRAM read GB | DRAM write GB | tensor pipe active % | SM GHz | achieved TFLOP/s* |")
    md.append("|---|---|---|---|---|---|---|")
    for i, r in enumerate(rows):
        t = ms(r, units)
        rd, wr = gb(r, units, "dram__bytes_read.sum"), gb(r, units, "dram__bytes_write.sum")
        # which layer: by bytes (2 GiB in / 2 GiB out = 128->128 @64; 1 GiB in = first conv on 1024 images)
        kind, flop = "?", float("nan")
        if rd > 1.9 and wr > 1.9:
            kind, flop = "128→128 @64² (plain)", 2048 * 1.20796e9
        elif rd > 1.9 and wr < 0.7:
            kind, flop = "128→128 @64² + MaxPool", 2048 * 1.20796e9
        elif rd < 1.3 and wr > 1.9:
            kind, flop = "128→128 @64², 1024 images fanned out ×2 (init_conv.conv2)", 1024 * 1.20796e9
        elif 0.4 < rd < 0.7 and wr > 0.9:
            kind, flop = "128→256 @32²", 2048 * 0.60398e9
        elif 0.9 < rd < 1.3 and 0.9 < wr < 1.3:
            kind, flop = "256→256 @32²", 2048 * 1.20796e9
        elif 0.9 < rd < 1.3 and wr < 0.4:
            kind, flop = "256→256 @32² + MaxPool", 2048 * 1.20796e9
        md.append(f"| {37 + i}: {kind} | {t:.3f} | {rd:.3f} | {wr:.3f} | "
                  f"{fnum(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                  f"{fnum(r, 'sm__cycles_elapsed.avg.per_second'):.2f} | {flop / t / 1e9:.0f} |")
    md.append("\n*under ncu (serialised, its own clocks). DRAM traffic equals the algorithmic bytes of each layer (NHWC bf16 in "
              "+ out): the halo tile and its nine tap views are served from shared memory, the weights from L2.\n")
    bres = os.path.join(O, "ev_prof_gemm_bres.ncu-rep")
    if os.path.exists(bres):
        rws, un = raw_rows(bres)
        md.append("`-k regex:gemm_bres -s 3 -c 3` on the same command (the resident-weight GEMMs of one step; up0 "
                  "`[2048,256]x[256,65536]`, up1 and up2 = the 2x2 transposed convolutions with their pixel-shuffle store):\n")
        md.append("| launch | time us | DRAM read GB | DRAM write GB | DRAM GB/s | tensor pipe active % | grid |")
        md.append("|---|---|---|---|---|---|---|")
        for r in rws:
            t = ms(r, un)
            rd, wr = gb(r, un, "dram__bytes_read.sum"), gb(r, un, "dram__bytes_write.sum")
            kind = "up0" if wr < 0.4 else ("up1" if wr < 0.8 else "up2")
            md.append(f"| {kind} | {t * 1e3:.1f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / t * 1e3:.0f} | "
                      f"{fnum(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                      f"{fnum(r, 'launch__grid_size'):.0f} |")
        md.append("")
    for name, rep in (("conv_out_mma_kernel<32> (2048 images)", "ev_prof_conv_out.ncu-rep"),
                      ("conv_in_mma_kernel (1024 images)", "ev_prof_conv_in.ncu-rep"),
                      ("ddpm_step_kernel (1024 samples)", "ev_prof_ddpm.ncu-rep"),
                      ("gemm_tn9_kernel (3x3 weight gradient, training step at batch 256)", "ev_prof_wgrad.ncu-rep")):
        path = os.path.join(O, rep)
        if not os.path.exists(path):
            continue
        rws, un = raw_rows(path)
        r = rws[0]
        t = ms(r, un)
        rd, wr = gb(r, un, "dram__bytes_read.sum"), gb(r, un, "dram__bytes_write.sum")
        md.append(f"**{name}**: {t * 1e3:.1f} us, DRAM {rd:.3f} GB read + {wr:.3f} GB written = "
                  f"{(rd + wr) / t * 1e3:.0f} GB/s ({fnum(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % of "
                  f"peak per ncu), tensor pipe {fnum(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} %, "
                  f"issue slots {fnum(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %, "
                  f"{fnum(r, 'launch__registers_per_thread'):.0f} registers, grid {fnum(r, 'launch__grid_size'):.0f}.\n")
    md.append("SASS of `libcdm_b200.so` contains UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld), UTCBAR "
              "(tcgen05.commit) and HMMA (the mma.sync first/last convolutions): `cuobjdump -sass "
              "camels-diffusion-model_b200/libcdm_b200.so`.\n")
    if os.path.exists(os.path.join(O, "ev_train_launches.csv")):
        md.append("## 3. Launch list of one training step (batch 256, eager, `tools/train_breakdown.py 256`)\n")
        md.append(train_step_table(os.path.join(O, "ev_train_launches.csv")) + "\n")
    if os.path.exists(os.path.join(O, "ev_probe.log")):
        md.append("## 4. Kernel probes (`tools/gpu_probe.py`, plain runs)\n\n```")
        with open(os.path.join(O, "ev_probe.log")) as fh:
            md.append("".join(l for l in fh if l.startswith(("CASE", "   issuer"))).rstrip())
        md.append("```\n")
    md.append("The `.ncu-rep` files are kept out of git (4-11 MB each); regenerate with `tools/collect_evidence.sh`.\n")
    with open(os.path.join(P, f"{tag}_ncu_summary.md"), "w") as fh:
        fh.write("\n".join(md))
    print("wrote", os.path.join(P, f"{tag}_ncu_summary.md"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
