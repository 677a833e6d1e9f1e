"""Per-source-line hot spots of a kernel from an .ncu-rep (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [file_substring]
"""
import csv
import io
import subprocess
import sys


def main(rep, top=40, only=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, agg = "", None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
            d = dict(zip(hdr[4:], r[4:]))
            key = (fname, int(r[0]))
            e = agg.setdefault(key, {"src": r[1], "inst": 0, "smp": 0, "long_sb": 0, "barrier": 0, "short_sb": 0,
                                     "wait": 0, "mio": 0, "branch": 0})
            e["inst"] += int(d.get("Instructions Executed", 0) or 0)
            e["smp"] += int(d.get("# Samples", 0) or 0)
            for k, c in (("long_sb", "stall_long_sb"), ("barrier", "stall_barrier"), ("short_sb", "stall_short_sb"),
                         ("wait", "stall_wait"), ("mio", "stall_mio"), ("branch", "stall_branch_resolving")):
                e[k] += int(d.get(c, 0) or 0)
    ti = sum(e["inst"] for e in agg.values()) or 1
    ts = sum(e["smp"] for e in agg.values()) or 1
    print(f"total warp instructions {ti:.3e}, samples {ts}")
    items = [(k, e) for k, e in agg.items() if only is None or only in k[0]]
    items.sort(key=lambda ke: -ke[1]["smp"])
    print(" inst%  smp%  long barr shrt wait  mio  brch  file:line  source")
    for (f, ln), e in items[:top]:
        print(f"{100 * e['inst'] / ti:5.1f} {100 * e['smp'] / ts:5.1f} "
              f"{100 * e['long_sb'] / ts:5.1f} {100 * e['barrier'] / ts:4.1f} {100 * e['short_sb'] / ts:4.1f} "
              f"{100 * e['wait'] / ts:4.1f} {100 * e['mio'] / ts:4.1f} {100 * e['branch'] / ts:4.1f}  {f}:{ln}  {e['src'].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else None)
