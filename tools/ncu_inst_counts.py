"""Warp instructions and DRAM bytes per voxel-view update of the hot kernels AT BENCH SIZE, from an ncu CSV.

Capture (one GPU, under gpurun; bench.py itself is the workload, so the shapes are the headline's):

    ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\
smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,\
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none \
        -k regex:"walk_|sino_interleave" -c 12 --csv \
        --log-file gpurun_out/inst_counts.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu \
        --no-configs --no-view-block --solver-iters 0

    python tools/ncu_inst_counts.py gpurun_out/inst_counts.csv 1024 1024 > profiles/ncu_r02_bench_size.json

The first forward application (the launches before the first adjoint launch) and the first adjoint application (the
row-interleaving pass, when the plan uses it, plus the adjoint launch) are summed; `bench.py` divides these per-update
counts by launch durations it measures live (never by ncu's).  Wavefronts = l1tex data-pipe wavefronts (shared-memory
loads / stores / atomics and the few global accesses): the pipe retires one per clock and SM.
"""
import csv
import json
import sys


def main():
    path, n, views = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = {}
    order = []
    for r in rows:
        i = int(r["ID"])
        if i not in launches:
            launches[i] = {"name": r["Kernel Name"], "grid": r.get("Grid Size"), "block": r.get("Block Size")}
            order.append(i)
        launches[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    updates = float(n) ** 3 * views
    seq = [launches[i] for i in order]
    # first application of each direction: forward launches up to the first adjoint launch, split into
    # applications by the class launch count (the launches before the first adjoint are k whole applications)
    first_adj = next(k for k, l in enumerate(seq) if "adjoint" in l["name"])
    fwd_before = [l for l in seq[:first_adj] if "forward" in l["name"]]
    pre_adj = [l for l in seq[max(0, first_adj - 1):first_adj] if "interleave" in l["name"]]
    names = [l["name"] + "|" + str(l["grid"]) for l in fwd_before]
    per_app = next(k for k in range(1, len(names) + 1) if len(names) % k == 0 and names[:k] * (len(names) // k) == names)
    fwd = fwd_before[-per_app:]
    adj = pre_adj + [seq[first_adj]]

    def summarise(ls, key_name):
        inst = sum(l["smsp__inst_executed.sum"] for l in ls)
        dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in ls)
        dur = sum(l["gpu__time_duration.sum"] for l in ls)
        issue = sum(l["smsp__issue_active.avg.pct_of_peak_sustained_active"] * l["gpu__time_duration.sum"] for l in ls) / dur
        wf_key, wf_pct = "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
        wf = sum(l.get(wf_key, 0.0) for l in ls) if all(wf_key in l for l in ls) else None
        wfp = sum(l[wf_pct] * l["gpu__time_duration.sum"] for l in ls) / dur if all(wf_pct in l for l in ls) else None
        return {"kernel": key_name, "launches_per_application": len(ls), "warp_inst_per_application": inst,
                "warp_inst_per_update": inst / updates, "dram_bytes_per_application": dram,
                "dram_bytes_per_update": dram / updates, "issue_active_pct": issue,
                "l1tex_wavefronts_per_application": wf, "l1tex_wavefronts_per_update": wf / updates if wf else None,
                "l1tex_data_pipe_pct": wfp,
                "ncu_duration_ms_per_application_cold_serialised": dur / 1e6,
                "source": f"ncu smsp__inst_executed.sum / dram__bytes_*.sum over bench.py at {n}^3 x {views} views, one B200 "
                          f"(profiles/ncu_r02_bench_size.json)",
                "launches": [{"name": l["name"][:120], "grid": l["grid"], "warp_inst": l["smsp__inst_executed.sum"],
                              "l1tex_wavefronts": l.get(wf_key), "ms": l["gpu__time_duration.sum"] / 1e6} for l in ls]}

    out = {"shape": f"{n}^3 volume, {views} views, detector {n}x{n}", "updates_per_application": updates,
           "walk_forward_joint": summarise(fwd, fwd[0]["name"].split("<")[0].split("::")[-1]),
           "walk_adjoint": summarise(adj, adj[-1]["name"].split("<")[0].split("::")[-1])}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
