"""Summarise an `ncu --set full` report (run HERE, no GPU needed: `ncu -i x.ncu-rep --page raw --csv`).

    python tools/ncu_summarise.py gpurun_out/prof_r01.ncu-rep --scale-fwd F --scale-adj A --out profiles/ncu_summary.json

Prints a markdown table of the metrics DESIGN.md / bench.py quote and writes a JSON with the DRAM
traffic per launch.  `--scale-*` multiply the captured launch's traffic to the bench-size launch
(the capture runs a reduced problem: traffic scales with slices x views)."""
import argparse
import csv
import io
import json
import subprocess

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram read",
    "dram__bytes_write.sum": "dram write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram throughput %",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps active %",
    "sm__inst_executed.sum": "instructions (warp)",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue active %",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active": "LSU writeback %",
    "l1tex__data_pipe_lsu_wavefronts.sum": "LSU wavefronts",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1tex data pipe % (shared-memory wavefronts)",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared bank-conflict wavefronts",
    "lts__t_sector_hit_rate.pct": "L2 hit %",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "XU pipe %",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "LSU pipe %",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "FMA pipe %",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "ALU pipe %",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall math_pipe_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall lg_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall mio_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall barrier",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall not_selected",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--scale-fwd", type=float, default=1.0)
    ap.add_argument("--scale-adj", type=float, default=1.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col["Kernel Name"]
    out = {"note": args.note, "kernels": []}
    print("| kernel | " + " | ".join(KEYS.values()) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in data:
        name = r[kn].split("(")[0]
        vals = {}
        for k, label in KEYS.items():
            if k in col:
                v = r[col[k]].replace(",", "")
                try:
                    vals[label] = float(v)
                except ValueError:
                    vals[label] = v
                vals[label + " unit"] = units[col[k]]
        print(f"| {name[:60]} | " + " | ".join(f"{vals.get(l, '')}" for l in KEYS.values()) + " |")
        out["kernels"].append({"name": r[kn], **vals})

    def traffic(pattern, scale):
        ks = [k for k in out["kernels"] if pattern in k["name"]]
        if not ks:
            return None

        def tobytes(k, lab):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(k.get(lab + " unit", "byte"), 1)
            return float(k.get(lab, 0.0)) * mult

        return sum(tobytes(k, "dram read") + tobytes(k, "dram write") for k in ks) / len(ks) * scale

    out["forward_traffic_bytes_per_launch_at_bench_size"] = traffic("forward", args.scale_fwd)
    out["adjoint_traffic_bytes_per_launch_at_bench_size"] = traffic("adjoint", args.scale_adj)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
