"""profiles/roofline_traffic.json from an `ncu --set full` capture of bench.py's trace_kernel launch (read here, no GPU):
   python tools/update_roofline_traffic.py gpurun_out/prof_trace.ncu-rep "bench.py --steps 4 --warmup 3 ..."
The file is stamped with the sha of the kernel sources (bench.kernel_sources_sha), so that bench.py only reports these
numbers for the code they were measured on."""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

COUNTERS = {
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "duration_us_under_ncu": "gpu__time_duration.sum",
}


def num(s):
    return float(s.replace(",", ""))


def main():
    rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    r = next(r for r in rows[2:] if "trace_kernel" in r[col["Kernel Name"]])

    def bytes_of(name):
        v, u = num(r[col[name]]), units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    traffic = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    counters = {k: num(r[col[m]]) for k, m in COUNTERS.items() if m in col and r[col[m]] != ""}
    if "duration_us_under_ncu" in counters and units[col[COUNTERS["duration_us_under_ncu"]]].lower().startswith("ns"):
        counters["duration_us_under_ncu"] /= 1e3
    d = {"trace_kernel_dram_bytes_per_launch": int(traffic), "counters": counters, "kernel": r[col["Kernel Name"]],
         "kernel_sources_sha": bench.kernel_sources_sha(),
         "source": f"{Path(rep).name}: ncu --set full --clock-control none of `{cmd}`; dram__bytes_read.sum + dram__bytes_write.sum "
                   "of one C4 re-trace launch; counters from the same launch (cold cache, serialised: shares, not absolute times)"}
    (ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(d, indent=1) + "\n")
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
