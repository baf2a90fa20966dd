"""Summarise an ncu --csv launch list (one row per launch and metric) per kernel family.
    python tools/ncu_summarize.py gpurun_out/launches.csv [--json out.json]"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    per = defaultdict(dict)
    names = {}
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) if r[iv] not in ("", "n/a") else 0.0
        unit = r[iu]
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)   # -> ms
        if unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        per[r[iid]][r[im]] = v
        names[r[iid]] = re.sub(r"<.*|\(.*", "", r[ik]).replace("void ", "").replace("cavit::", "")
    agg = defaultdict(lambda: defaultdict(float))
    for i, m in per.items():
        a = agg[names[i]]
        a["n"] += 1
        for k, v in m.items():
            a[k] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    out = {"total_ms": tot, "launches": int(sum(a["n"] for a in agg.values())), "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        d = {"n": int(a["n"]), "ms": round(a["gpu__time_duration.sum"], 4), "share": round(a["gpu__time_duration.sum"] / tot, 4)}
        if "dram__bytes_read.sum" in a:
            d["dram_read_bytes"] = a["dram__bytes_read.sum"]
            d["dram_write_bytes"] = a["dram__bytes_write.sum"]
            d["traffic_bytes_per_launch"] = (a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / a["n"]
        for mk in a:
            if "pct" in mk:
                d[mk + ".mean"] = round(a[mk] / a["n"], 3)
        out["kernels"][k] = d
        print(f"{k:44s} n={d['n']:4d} {d['ms']:9.3f} ms {100 * d['share']:5.1f}%" +
              (f"  dram {(d['dram_read_bytes'] + d['dram_write_bytes']) / 1e9:7.3f} GB" if "dram_read_bytes" in d else ""))
    print(f"total {tot:.3f} ms over {out['launches']} launches")
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
