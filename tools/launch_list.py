"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`) per kernel.

    python tools/launch_list.py gpurun_out/launches.csv profiles/name.json "the command that was profiled"
"""
import csv
import json
import sys


def main(src, out, command=""):
    rows = [r for r in csv.reader(open(src, newline="")) if r]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    col = {n: i for i, n in enumerate(rows[hdr])}
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = per.setdefault(r[col["Kernel Name"]], {"launches": 0, "ms": 0.0})
        k["launches"] += 1
        k["ms"] += ms
    total = sum(k["ms"] for k in per.values())
    for k in per.values():
        k["share"] = k["ms"] / total if total else 0.0
        k["ms_per_launch"] = k["ms"] / k["launches"]
    per = dict(sorted(per.items(), key=lambda kv: -kv[1]["ms"]))
    json.dump({"command": command, "total_ms": total, "kernels": per}, open(out, "w"), indent=1)
    for name, k in list(per.items())[:8]:
        print(f"{k['share']:7.3f} {k['launches']:6d} {k['ms_per_launch']:9.4f} ms  {name[:90]}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
