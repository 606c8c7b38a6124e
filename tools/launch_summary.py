"""Reduce an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>)
to per-kernel counts, total time and share.
    python tools/launch_summary.py gpurun_out/launches.csv profiles/out_summary.json "description" """
import collections
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)                         # drop the argument list
    name = name.replace("void ", "").replace("gpe::", "").replace("<unnamed>::", "")
    return name


def main():
    src, out, desc = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        t_ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        key = short(r["Kernel Name"])
        a = agg.setdefault(key, {"launches": 0, "ms": 0.0, "grids": collections.Counter()})
        a["launches"] += 1
        a["ms"] += t_ms
        a["grids"][r["Grid Size"]] += 1
        total += t_ms
        n += 1
    rows = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        rows.append({"kernel": k, "launches": a["launches"], "ms": round(a["ms"], 3), "share": round(a["ms"] / total, 4),
                     "top_grids": dict(a["grids"].most_common(4))})
    json.dump({"description": desc, "source": src, "launches": n, "total_ms": round(total, 3), "kernels": rows},
              open(out, "w"), indent=1)
    for r in rows[:14]:
        print("%-60s %6d launches %10.2f ms %6.2f %%" % (r["kernel"][:60], r["launches"], r["ms"], 100 * r["share"]))
    print("total %.1f ms over %d launches" % (total, n))


if __name__ == "__main__":
    main()
