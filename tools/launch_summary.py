"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X): per kernel the number of
launches, total and mean device time and the share of the total.  usage: python tools/launch_summary.py X.csv "cmd" """
import csv
import sys
from collections import OrderedDict


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {h: i for i, h in enumerate(rows[hdr])}
    acc = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0]
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in acc.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none")
    if cmd:
        print(f"#   {cmd}")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stages_ms_per_step, not absolutes")
    for name, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} n={n:4d} total={us:10.1f} us share={100 * us / tot:5.1f}% avg={us / n:8.1f} us")


if __name__ == "__main__":
    main()
