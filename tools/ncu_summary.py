"""Turn an `ncu --set full` capture of one kernel launch into the summary kept under profiles/ and (optionally) the
per-unit DRAM traffic entry of profiles/ncu_traffic.json that bench.py scales to the launch it times.

    python tools/ncu_summary.py gpurun_out/prof_beam.ncu-rep profiles/ncu_beam_r02_summary.txt \
        --traffic beam_kernel --units 177600 --unit-name bytes_per_utterance_frame --note "592 utterances x 300 frames"
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--cmd", default="")
    ap.add_argument("--traffic", default="", help="kernel key in profiles/ncu_traffic.json to update")
    ap.add_argument("--units", type=float, default=0.0); ap.add_argument("--unit-name", default="bytes_per_utterance_frame")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(hdr)}
    lines = [f"# {a.cmd}" if a.cmd else "# ncu --set full --clock-control none --import-source on, one launch", f"# {a.note}",
             f"Kernel Name [] = {vals[col['Kernel Name']]}"]
    got = {}
    for k in WANT:
        if k in col:
            lines.append(f"{k} [{units[col[k]]}] = {vals[col[k]]}")
            got[k] = (float(vals[col[k]].replace(",", "")), units[col[k]])
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if a.traffic and a.units > 0:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = sum(got[k][0] * scale[got[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        import bench
        tfile = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        tj = json.load(open(tfile)) if os.path.exists(tfile) else {}
        tj[a.traffic] = {a.unit_name: tot / a.units, "capture": os.path.relpath(a.out, ROOT) + (f" ({a.note})" if a.note else ""),
                         "source_sha": bench.kernel_source_sha()}
        json.dump(tj, open(tfile, "w"), indent=1)
        print(f"{a.traffic}: {tot / a.units:.0f} bytes per unit, source_sha {tj[a.traffic]['source_sha']}")


if __name__ == "__main__":
    main()
