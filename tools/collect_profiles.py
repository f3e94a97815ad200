"""gpurun_out/<tag>_* (written by tools/gpu_round.sh on the GPU box) -> the committed evidence under profiles/.

    python tools/collect_profiles.py r01d [layer,layer,...]

  profiles/<tag>_bench.json, _layer_times_<mode>.txt, _pytest_gpu.log, _smoke.log      copies
  profiles/<tag>_ncu_full_<mode>.csv        one row per profiled launch: the metrics the roofline discussion uses (ncu --set full)
  profiles/<tag>_ncu_source_top_<mode>_<k>.txt  SASS lines of launch k by warp-state samples (ncu --page source)
  profiles/<tag>_ncu_launches.csv.gz + _summary.txt   ncu launch list of bench.py (gpu__time_duration.sum per launch), share per kernel
  profiles/<tag>_ncu_traffic_<mode>.csv.gz + profiles/traffic_<mode>.json   DRAM bytes of every launch of one forward (read by bench.py)
"""
import collections
import csv
import gzip
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
DEFAULT_LAYERS = ("resblocks.2.convs1.2,resblocks.3.convs1.0,resblocks.5.convs1.2,resblocks.5.convs2.2,resblocks.8.convs1.2,resblocks.6.pair.0,"
                  "resblocks.7.pair.1,resblocks.9.pair.0,resblocks.10.pair.1,resblocks.11.pair.2,ups.1,conv_post")
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def short(name):
    return name.split("(")[0].replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "").replace("hfg::", "").strip()


def full_summary(tag, mode):
    src = os.path.join(OUT, f"{tag}_prof_{mode}_raw.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    plain = os.path.join(OUT, f"{tag}_prof_plain_{mode}.log")
    with open(os.path.join(PROF, f"{tag}_ncu_full_{mode}.csv"), "w", newline="") as f:
        f.write(f'"# ncu --set full --clock-control none --import-source on --profile-from-start off; HFG_NCU_LAYERS brackets the listed layers; '
                f'python tools/layer_times.py --mode {mode} --B 16 --T 862 --reps 1 --warm 1 (two forwards: first half of the launches = warm-up)"\n')
        w = csv.writer(f)
        w.writerow(["launch"] + [f"{n} [{units[i]}]" if units[i] else n for n, i in cols])
        for j, r in enumerate(data):
            vals = [r[i] for _, i in cols]
            vals[0] = short(vals[0])
            w.writerow([j] + vals)


def source_top(tag, mode, k, n=40):
    src = os.path.join(OUT, f"{tag}_prof_{mode}_source_launch{k}.csv")
    if not os.path.exists(src) or os.path.getsize(src) < 1000:
        return
    rows = list(csv.reader(open(src)))
    kernel = short(rows[0][1]) if len(rows[0]) > 1 else "?"
    hdr = rows[1]
    data = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(d["# Samples"] or 0) for d in data)
    top = sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:n]
    with open(os.path.join(PROF, f"{tag}_ncu_source_top_{mode}_launch{k}.txt"), "w") as f:
        f.write(f"# ncu --page source, profiled launch {k} ({kernel}), {mode}: SASS lines by warp-state samples (total {tot}); top stall reasons per line\n")
        for d in top:
            st = sorted(((s, int(d[s] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
            s = int(d["# Samples"] or 0)
            f.write(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}%  {d['Source'][:72]:72s} {st[0][0]}={st[0][1]} {st[1][0]}={st[1][1]}\n")


def launches(tag):
    src = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_ncu_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 900: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary (bf16x3)\n")
        f.write("# kernel, launches, total_us, share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k}, {a[0]}, {a[1]:.1f}, {a[1] / tot:.4f}\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(PROF, f"{tag}_ncu_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)


def traffic(tag, mode):
    src = os.path.join(OUT, f"{tag}_traffic_{mode}.csv")
    if not os.path.exists(src):
        return
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "traffic_summary.py"), src, mode, os.path.join(PROF, f"traffic_{mode}.json")],
                   check=True, stdout=subprocess.DEVNULL)
    with open(src, "rb") as fi, gzip.open(os.path.join(PROF, f"{tag}_ncu_traffic_{mode}.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    for name in ("bench.json", "layer_times_bf16.txt", "layer_times_bf16x3.txt", "pytest_gpu.log", "smoke.log"):
        src = os.path.join(OUT, f"{tag}_{name}")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PROF, f"{tag}_{name}"))
    for mode in ("bf16", "bf16x3"):
        full_summary(tag, mode)
        for k in range(12):
            source_top(tag, mode, k)
        traffic(tag, mode)
    launches(tag)
    print("\n".join(sorted(f for f in os.listdir(PROF) if f.startswith(tag) or f.startswith("traffic_"))))


if __name__ == "__main__":
    main()
