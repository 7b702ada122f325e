"""Hot source lines of a profiled kernel: python tools/ncu_source_lines.py <rep> [top]  (reads the .ncu-rep here;
needs -lineinfo and --import-source on).  Per CUDA source line: share of stall samples, share of executed warp
instructions, average active lanes."""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    hdr, cur, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            print("#", r[1][:150])
        elif len(r) >= 2 and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].strip().isdigit():
            ti = hdr.index("Thread Instructions Executed")
            agg.append((num(r[6]), num(r[7]), num(r[ti]), cur, int(r[0]), r[1].strip()[:110]))
    ts, ti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    print(f"# {ts} stall samples, {ti} warp instructions; per line: samples %, instructions %, active lanes")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{a[0] * 100 / ts:5.1f}% {a[1] * 100 / ti:5.1f}% {a[2] / max(a[1], 1):5.1f}  {a[3]}:{a[4]}  {a[5]}")


if __name__ == "__main__":
    main()
