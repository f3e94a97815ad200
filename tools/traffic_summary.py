"""ncu metrics-only CSV (dram bytes per launch of one forward) -> profiles/traffic_<mode>.json, read by bench.py.
    python tools/traffic_summary.py gpurun_out/r01c_traffic_bf16x3.csv bf16x3 profiles/traffic_bf16x3.json
"""
import collections
import csv
import json
import sys


def main():
    src, mode, out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, mi, ui, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9}
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[start + 1:]:
        names[r[idi]] = r[ki].split("(")[0].replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "").strip()
        per[r[idi]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    ids = sorted(per, key=int)
    half = len(ids) // 2          # two forwards were captured (warm-up, timed): keep the second
    agg = collections.defaultdict(lambda: {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "seconds": 0.0})
    for i in ids[half:]:
        k = names[i].split("<")[0]
        a = agg[k]
        a["launches"] += 1
        a["dram_read_bytes"] += per[i].get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += per[i].get("dram__bytes_write.sum", 0.0)
        a["seconds"] += per[i].get("gpu__time_duration.sum", 0.0)
    res = {"mode": mode, "workload": "V1, 16 x 862-frame mels, one forward", "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (metrics-only pass over every launch)",
           "kernels": {}}
    for k, a in agg.items():
        res["kernels"][k] = {"launches": a["launches"], "traffic_bytes_per_launch": (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"],
                             "dram_read_bytes": a["dram_read_bytes"], "dram_write_bytes": a["dram_write_bytes"], "ncu_seconds": a["seconds"]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
