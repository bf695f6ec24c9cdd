#!/usr/bin/env python
"""Headline benchmark: XRayTransform3D forward + adjoint pair, voxel-view updates/s.

Workload (BASELINE.json configs[4] operator / north_star target): 3D parallel-beam
XRayTransform3D, 1024^3 volume, 1024 views about axis 0, 1024x1024 detector.  One step = one
forward projection + one back projection on DENSE input (standard-normal volume, its sinogram); the
tanglecube phantom (56 % zeros, for which the forward skips all-zero flushes) is timed beside it as
`phantom`.  With N GPUs the operator is partitioned into N z-slabs (volume slices <-> detector rows),
one process per GPU, no data-path collective: total work is fixed, so scaling is "strong".  The
partition that DOES communicate (view blocks: C3 = BASELINE.json configs[2] and a tilted 3D geometry,
back projection + NCCL reduce-scatter or the exchange fused into the kernel over NVLink peer memory)
is timed at every N as `view_block`.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 3 --warmup 3
    python bench.py --impl reference         # the CPU arm (oracle C port, all host threads)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "XRayTransform3D fwd+adj voxel-view updates/s"
UNIT = "voxel-view updates/s"


def workload(args):
    n, v = args.size, args.views
    return dict(N=(n, n, n), D=(n, n), V=v)


def cpu_sample(wl, slices=16, views=64):
    """Bounded sample of the same workload for the CPU arm: a z-slab of `slices` slices and the
    first `views` of the V views (same plane size, same geometry family)."""
    n = wl["N"][1]
    s = min(slices, wl["N"][0])
    v = min(views, wl["V"])
    return dict(N=(s, n, n), D=(s, n), V=v, desc=f"z-slab of {s} slices x first {v} of {wl['V']} views of the {n}^3 workload, fwd+adj")


def time_cpu_port(sample, reps=1):
    """Time the oracle's C port (OpenMP, all host threads) on the sample.  Only this function and
    the parity check in smoke()/tests touch oracle/."""
    from oracle import xray_c as C
    from oracle import xray_np as O

    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports
    # OMP_NUM_THREADS=1 to every rank, which would time a single core)
    C.set_num_threads(len(os.sched_getaffinity(0)))
    N, D, V = sample["N"], sample["D"], sample["V"]
    ang = np.linspace(0, np.pi, sample.get("V_total", V), endpoint=False)[:V, None]
    M = O.matrices_from_euler_angles(N, D, "X", ang).astype(np.float32)
    # centre the slab on the rotation axis like the full volume (rows only select the slice)
    x = np.random.default_rng(0).random(N, dtype=np.float32)
    C.project_3d(x[:1], M[:1], (1, D[1]), fused=True)  # warm up the thread pool
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        y = C.project_3d(x, M, D, fused=True)
        C.back_project_3d(y, M, N)
        best = min(best, time.perf_counter() - t0)
    updates = 2.0 * float(np.prod(N)) * V
    return updates / best, best, C.num_threads()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.1):
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pin_to_gpu_numa_node(torch, local):
    """Run this rank (and first-touch its page-locked buffers) on the NUMA node its GPU hangs off: with 8 ranks
    all on node 0, half of the host<->device traffic crosses the socket interconnect.  sysfs only (no libnuma):
    /sys/bus/pci/devices/<gpu>/numa_node -> /sys/devices/system/node/nodeK/cpulist, intersected with the CPUs
    this process may use.  Returns a one-line note for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            q = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                               capture_output=True, text=True, timeout=10).stdout.strip()
            bus = q[-12:].lower() if q else None  # 00000000:1B:00.0 -> 0000:1b:00.0
        if not bus:
            return "GPU PCI bus id unavailable: not pinned"
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(path):
            return f"{path} missing: not pinned"
        node = int(open(path).read().strip())
        if node < 0:
            return "GPU reports no NUMA node: not pinned"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return f"GPU on NUMA node {node}, none of its CPUs allowed for this process: not pinned"
        os.sched_setaffinity(0, use)
        return f"pinned to NUMA node {node} ({len(use)} of {len(allowed)} allowed CPUs)"
    except Exception as e:  # sysfs layout / permissions: measurement goes on unpinned
        return f"not pinned ({type(e).__name__}: {e})"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_profile():
    """Warp instructions and DRAM bytes per voxel-view update of the two hot kernels, from the committed ncu
    capture AT BENCH SIZE (profiles/ncu_r02_bench_size.json, written by tools/ncu_inst_counts.py from
    `ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum` over this very script).
    Instruction counts are a property of the code and the shape, not of the data or the clocks; the launch
    durations they are divided by are measured live, below."""
    p = os.path.join(ROOT, "profiles", "ncu_r02_bench_size.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


def other_configs(torch, sb, dev):
    """Forward / adjoint timings of BASELINE.json configs[1..3] on one GPU (CUDA events, 3 warm-ups;
    C2's working set is smaller than L2, so L2 is flushed before every timed call)."""

    def flush_l2():
        torch.empty(64 * 1024 * 1024, device=dev).fill_(0.0)  # 256 MB > 126 MB L2

    def time_op(A, updates, reps, flush):
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(A.input_shape, device=dev, generator=g)
        y = torch.randn(A.output_shape, device=dev, generator=g)
        res = {}
        for tag, fn, arg in (("fwd", A.__call__, x), ("adj", A.adj, y)):
            for _ in range(3):
                fn(arg)
            ts = []
            for _ in range(reps):
                if flush:
                    flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(arg)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[tag + "_ms"] = float(np.mean(ts))
        Ax, ATy = A(x), A.adj(y)
        gap = abs(torch.sum(Ax.double() * y.double()).item() - torch.sum(x.double() * ATy.double()).item()) / (
            torch.linalg.vector_norm(Ax.double()).item() * torch.linalg.vector_norm(y.double()).item())
        info = A.plan_info()
        res.update({"pair_updates_per_s": 2 * updates / (res["fwd_ms"] + res["adj_ms"]) * 1e3, "adjoint_gap": gap,
                    "kernel_path": info["path_name"], "l2": "flushed before every call" if flush else "working set > L2"})
        return res

    out = {}
    A = sb.XRayTransform2D((512, 512), np.linspace(0, np.pi, 360, endpoint=False))
    out["C2 2D 512^2 x 360 views, 725 bins"] = time_op(A, 512 * 512 * 360, 10, True)
    A = sb.XRayTransform2D((4096, 4096), np.linspace(0, np.pi, 2048, endpoint=False))
    out["C3 2D 4096^2 x 2048 views, 5793 bins (one GPU)"] = time_op(A, 4096 * 4096 * 2048, 3, False)
    n, V = 512, 720
    M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sb.XRayTransform3D((n,) * 3, M, (n, n))
    out["C4 3D 512^3 x 720 views, det 512^2 (one GPU)"] = time_op(A, n ** 3 * V, 3, False)
    return out


def view_block_configs(torch, dist, sb, sharded, dev, world, barrier, reduce_max):
    """The partition of the path that has a real exchange step, timed at this N (BASELINE.json
    configs[2]): contiguous view blocks per rank.  C3 (2D 4096^2 x 2048 views): image replicated, partial
    back projections summed row block by row block into their owners -- by ONE NCCL reduce_scatter
    ("nccl") or inside the back-projection kernel's epilogue over NVLink peer memory ("peer",
    xct_adjoint_scatter + sharded.PeerBlocks).  Tilted 3D (256^3 x 64 views, 74 degree XY tilt, general
    matrices): all-gather of the slab-sharded volume before the forward, per-slab reduce / fused exchange after
    the adjoint.  CUDA events between barriers, 3 warm-ups, max over ranks."""

    def timeit(fn, reps=3):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1) / reps)

    out = {}
    g = torch.Generator(device=dev).manual_seed(1234)  # same seed on every rank: replicated inputs agree
    n, V = 4096, 2048
    angles = np.linspace(0, np.pi, V, endpoint=False)
    A = sharded.ViewShardedXRayTransform2D((n, n), angles)
    x = torch.randn((n, n), device=dev, generator=g)
    y = A.project(x)
    f_ms = timeit(lambda: A.project(x))
    a_nccl = timeit(lambda: A.back_project(y, scatter=True))
    a_local = timeit(lambda: A.local.back_project(y))
    a_peer = None
    if world > 1:
        P = sharded.ViewShardedXRayTransform2D((n, n), angles, exchange="peer")
        a_peer = timeit(lambda: P.back_project(y))
        ref, got = A.back_project(y), P.back_project(y)
        peer_err = reduce_max((torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref)).item())
        P.close()
        del P, ref, got
    upd = float(n) * n * V
    best = a_nccl if a_peer is None else min(a_nccl, a_peer)
    out["C3 2D 4096^2 x 2048 views"] = {
        "views_per_rank": A.views[1] - A.views[0], "fwd_ms": f_ms, "adj_ms_nccl_reduce_scatter": a_nccl,
        "adj_ms_fused_peer_exchange": a_peer, "adj_ms_kernel_only": a_local,
        "pair_updates_per_s_nccl": 2 * upd / (f_ms + a_nccl) * 1e3, "pair_updates_per_s_best": 2 * upd / (f_ms + best) * 1e3,
        "peer_vs_nccl_rel_l2": None if a_peer is None else peer_err,
        "collective": "none (one rank)" if world == 1 else
                      "ncclReduceScatter(sum) of the 64 MB partial images over equal row blocks | fused: stores into the "
                      "owners' slots over NVLink + slot sum"}
    del A, x, y
    torch.cuda.empty_cache()

    n, V = 256, 64
    angs = np.stack([np.linspace(0, np.pi, V, endpoint=False), np.full(V, np.deg2rad(74.0))], 1)
    D = (n + 64, n + 64)
    M = sb.matrices_from_euler_angles((n,) * 3, D, "XY", angs)
    A = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D)
    xs = torch.randn(A.local_input_shape, device=dev, generator=g)
    ys = A.project(xs)
    f_ms = timeit(lambda: A.project(xs))
    a_nccl = timeit(lambda: A.back_project(ys))
    a_peer = None
    if world > 1:
        P = sharded.ViewShardedXRayTransform3D((n,) * 3, M, D, exchange="peer")
        a_peer = timeit(lambda: P.back_project(ys))
        P.close()
        del P
    upd = float(n) ** 3 * V
    best = a_nccl if a_peer is None else min(a_nccl, a_peer)
    info = A.full.plan_info(torch.cuda.current_device()) if A.full is not None else {}
    out["3D 256^3 x 64 views, XY tilt 74 deg (general matrices)"] = {
        "fwd_ms": f_ms, "adj_ms_nccl_reduce": a_nccl, "adj_ms_fused_peer_exchange": a_peer,
        "pair_updates_per_s_nccl": 2 * upd / (f_ms + a_nccl) * 1e3, "pair_updates_per_s_best": 2 * upd / (f_ms + best) * 1e3,
        "kernel_path": info.get("path_name"),
        "collective": "none (one rank)" if world == 1 else
                      "forward: ncclAllGather of the slab-sharded volume; adjoint: per-slab ncclReduce overlapped with the "
                      "next slab's kernel | fused peer exchange"}
    del A, xs, ys
    torch.cuda.empty_cache()

    # separable 3D geometry in VIEW-BLOCK mode (C4's operator: 512^3, 720 views, det 512^2): what z-slab sharding
    # does without any exchange, done the communicating way -- the routed epilogue of the walk adjoint
    n, V = 512, 720
    M = sb.matrices_from_euler_angles((n,) * 3, (n, n), "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    A = sharded.ViewShardedXRayTransform3D((n,) * 3, M, (n, n))
    xs = torch.randn(A.local_input_shape, device=dev, generator=g)
    ys = A.project(xs)
    f_ms = timeit(lambda: A.project(xs))
    a_nccl = timeit(lambda: A.back_project(ys))
    a_peer = None
    if world > 1:
        P = sharded.ViewShardedXRayTransform3D((n,) * 3, M, (n, n), exchange="peer")
        a_peer = timeit(lambda: P.back_project(ys))
        P.close()
        del P
    upd = float(n) ** 3 * V
    best = a_nccl if a_peer is None else min(a_nccl, a_peer)
    out["C4 3D 512^3 x 720 views in view blocks (separable geometry, walk kernels)"] = {
        "fwd_ms": f_ms, "adj_ms_nccl_reduce": a_nccl, "adj_ms_fused_peer_exchange": a_peer,
        "pair_updates_per_s_nccl": 2 * upd / (f_ms + a_nccl) * 1e3, "pair_updates_per_s_best": 2 * upd / (f_ms + best) * 1e3,
        "collective": "none (one rank)" if world == 1 else
                      "forward: ncclAllGather of the volume slabs (512 MB); adjoint: per-slab ncclReduce | fused: the walk adjoint's "
                      "routed epilogue stores into the owners' slots over NVLink, flag rendezvous, slot sum"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args)
    sample = cpu_sample(wl)
    sample["V_total"] = wl["V"]
    vals = []
    for _ in range(max(0, args.warmup if args.warmup < 2 else 1)):
        time_cpu_port(sample)
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        v, sec, threads = time_cpu_port(sample)
        vals.append((v, sec))
    total = time.perf_counter() - t_all0
    value = float(np.mean([v for v, _ in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"XRayTransform3D {wl['N'][0]}^3 x {wl['V']} views, det {wl['D'][0]}x{wl['D'][1]}, fwd+adj",
                   "note": "reference arm = oracle C port of scico/linop/xray/_xray3d.py (JAX is not installable here), "
                           "each step a bounded sample normalised to updates/s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample["desc"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="volume edge (default: the 1024^3 headline workload)")
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C4 timings")
    ap.add_argument("--no-view-block", action="store_true", help="skip the view-block (communicating) partition timings")
    ap.add_argument("--solver-iters", type=int, default=3, help="TV-PDHG iterations timed after the operator bench (0 = skip)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import scico_b200 as sb
    from scico_b200 import _lib, geometry, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- scico_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa_note = pin_to_gpu_numa_node(torch, local)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver counts ranks from NCCL's own log); whatever NCCL
        # prints are separate lines, the JSON line below is printed once, by rank 0, at the very end
        dist.init_process_group("nccl", device_id=torch.device(dev))

    wl = workload(args)
    N, D, V = wl["N"], wl["D"], wl["V"]
    M = sb.matrices_from_euler_angles(N, D, "X", np.linspace(0, np.pi, V, endpoint=False)[:, None])
    assert geometry.is_axis0_separable(M)
    # z-slab of this rank: slices [z0, z1) <-> detector rows [r0, r1); no data-path collective
    SA = sharded.SlabShardedXRayTransform3D(N, M, D, rank=rank, world_size=world)
    assert not SA.halo_rows(), "bench geometry has aligned rows: no halo exchange"
    (z0, z1), (r0, r1) = SA.slab, SA.rows
    A = SA.local  # the rank-local XRayTransform3D (slice_offset = z0, detector rows [r0, r1))
    info = A.plan_info(local)

    # synthetic phantom: tanglecube (scico/examples.py:529-581) evaluated on the device
    def tangle(z_lo, z_hi):
        lin = lambda n: torch.linspace(-1.0, 1.0, n, device=dev, dtype=torch.float32) * 3.0
        zz = lin(N[0])[z_lo:z_hi, None, None]
        yy, xx = lin(N[1])[None, :, None], lin(N[2])[None, None, :]
        val = (xx**4 - 5 * xx**2 + yy**4 - 5 * yy**2 + zz**4 - 5 * zz**2 + 11.8) * 0.2 + 0.5
        return torch.where(val <= 2.0, 2.0 - val, torch.zeros_like(val)).clamp_(min=0.0).contiguous()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(val):
        if world == 1:
            return val
        t = torch.tensor([val], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def time_pairs(x, y, warmup, steps):
        """`steps` forward + adjoint pairs on resident inputs: (total ms, fwd ms, adj ms), max over ranks."""
        for _ in range(warmup):
            A(x)
            A.adj(y)
        barrier()
        e0, e1 = ev(), ev()
        fe = [(ev(), ev(), ev()) for _ in range(steps)]
        e0.record()
        for k in range(steps):
            fe[k][0].record()
            A(x)
            fe[k][1].record()
            A.adj(y)
            fe[k][2].record()
        e1.record()
        barrier()
        return (reduce_max(e0.elapsed_time(e1)), reduce_max(float(np.mean([a.elapsed_time(b) for a, b, _ in fe]))),
                reduce_max(float(np.mean([b.elapsed_time(c) for _, b, c in fe]))))

    # headline input: DENSE (standard-normal volume, no zeros: nothing for the forward's zero-block flush
    # predication to skip) and its own sinogram
    x = torch.randn((z1 - z0,) + N[1:], device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
    y = A(x)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        A(x)
        A.adj(y)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.launch_count_reset()
    t_wall0 = time.perf_counter()
    total_ms, fwd_ms, adj_ms = time_pairs(x, y, 0, args.steps)
    t_wall1 = time.perf_counter()
    launches = _lib.launch_count()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # the same step on the tanglecube phantom (the reference examples' test object; 56 % zeros)
    xp = tangle(z0, z1)
    yp = A(xp)
    p_total, p_fwd, p_adj = time_pairs(xp, yp, 1, args.steps)
    nz = reduce_max(float((xp != 0).float().mean().item()))
    del xp, yp

    ms_per_step = total_ms / args.steps
    updates_step = 2.0 * float(np.prod(N)) * V  # whole job: fwd + adj over the full volume
    value = updates_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (the forward: 60 % of a step) --------------------------------
    # The binding resources are ON-CHIP, not HBM: voxels (forward) / accumulators (adjoint) stay on the SM across
    # all views, so real DRAM traffic is ~0.3 % of what the per-view streaming model of SURVEY 8(d) assumes (that
    # model is kept below as `hbm_model`, labelled; its "fraction" exceeds 1).  Two on-chip ceilings are reported
    # per kernel and the higher fraction names the bound:
    #   issue : warp instructions per launch (ncu smsp__inst_executed.sum at THIS shape, committed under profiles/;
    #           a property of code + shape) / the launch duration measured live with CUDA events, against
    #           SMs x 4 schedulers x the SM clock sampled during the timed region (1 warp instruction per scheduler
    #           and cycle);
    #   shared: l1tex data-pipe wavefronts per launch (ncu l1tex__data_pipe_lsu_wavefronts.sum, same capture: the
    #           shared-memory loads / stores of the tile, the windows and the sinogram taps) / the same live
    #           duration, against SMs x 1 wavefront per cycle x the same clock.
    peak, peak_src = hbm_peak()
    loc_updates = float((z1 - z0) * N[1] * N[2]) * V
    loc_bytes = 4.0 * (loc_updates + float(V) * (r1 - r0) * D[1])
    kname = {0: "gen3d", 1: "plane", 2: "walk"}
    prof = ncu_profile()
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    if world > 1:
        t = torch.tensor([sm_mhz], device=dev, dtype=torch.float64)
        dist.broadcast(t, src=0)
        sm_mhz = float(t.item())
    issue_peak = n_sm * 4 * sm_mhz * 1e6 / 1e9  # G warp instructions / s
    wave_peak = n_sm * sm_mhz * 1e6 / 1e9       # G l1tex data-pipe wavefronts / s

    def kernel_roofline(key, ms, launches_per_app):
        k = prof.get(key) or {}
        ipu, wpu = k.get("warp_inst_per_update"), k.get("l1tex_wavefronts_per_update")
        inst = ipu * loc_updates if ipu else None
        wav = wpu * loc_updates if wpu else None
        ach_i = inst / (ms * 1e-3) / 1e9 if inst else None
        ach_w = wav / (ms * 1e-3) / 1e9 if wav else None
        issue = {"achieved": ach_i, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": ach_i / issue_peak if ach_i else None,
                 "warp_inst_per_launch": inst / launches_per_app if inst else None, "warp_inst_per_update": ipu,
                 "issue_active_pct_ncu": k.get("issue_active_pct")}
        shared = {"achieved": ach_w, "peak": wave_peak, "unit": "Gwavefront/s", "frac": ach_w / wave_peak if ach_w else None,
                  "wavefronts_per_launch": wav / launches_per_app if wav else None, "wavefronts_per_update": wpu,
                  "data_pipe_busy_pct_ncu": k.get("l1tex_data_pipe_pct")}
        top, bound = issue, "issue"
        if (shared["frac"] or 0.0) > (issue["frac"] or 0.0):
            top, bound = shared, "shared"
        return {"bound": bound, "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                "issue": issue, "shared": shared,
                "launch_ms": ms / launches_per_app, "launches_per_application": launches_per_app,
                "traffic": (k.get("dram_bytes_per_update") * loc_updates / launches_per_app) if k.get("dram_bytes_per_update") else None,
                "traffic_source": k.get("source"),
                "hbm_model": {"achieved": loc_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": loc_bytes / (ms * 1e-3) / 1e9 / peak,
                              "note": "SURVEY 8(d) per-view streaming model (4 B per update + the sinogram once): NOT a lower "
                                      "bound for kernels that keep voxels in registers across views; kept for reference"}}

    an = A.analyse()
    n_adj_launch = 2 if an.get("adj_interleaved") else 1
    n_fwd_launch = max(1, int(launches) // max(1, args.steps) - n_adj_launch)  # per step: forward class launches + the adjoint's
    fwd_name = (f"{'walk_forward_tile_kernel' if an.get('fwd_tile') else 'walk_forward_joint_kernel'}<Geom3> "
                f"({n_fwd_launch} class launches per application)" if info.get("fwd_joint")
                else f"{kname[info['fwd_kernel']]}_forward_kernel<Geom3> ({n_fwd_launch} launches per application, one per view class)")
    roofline = {"kernel": fwd_name, "sm_clock_mhz_in_timed_region": sm_mhz, "sms": n_sm,
                "peak_source": "issue: SMs x 4 warp schedulers x SM clock; shared: SMs x 1 l1tex data-pipe wavefront per cycle x SM "
                               "clock; the clock is what nvidia-smi reported during the timed region",
                "hbm_peak_source": peak_src + ", of measured"}
    roofline.update(kernel_roofline("walk_forward_joint", fwd_ms, n_fwd_launch))
    adj_name = ("walk_adjoint_vec_kernel<Geom3> (row-interleaving pass + 1 launch per application, TMA-staged window of the "
                "slice-interleaved sinogram)" if an.get("adj_interleaved")
                else f"{kname[info['adj_kernel']]}_adjoint_kernel<Geom3> (1 launch per application"
                + (", TMA-staged sinogram window)" if info.get("adj_tma") else ")"))
    roofline["adjoint"] = {"kernel": adj_name}
    roofline["adjoint"].update(kernel_roofline("walk_adjoint", adj_ms, n_adj_launch))

    # end-to-end through the public API with HOST buffers (pinned): H2D + kernels + D2H per call
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        sh = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
        sh.copy_(y)
        xh_np, sh_np = xh.numpy(), sh.numpy()
        # page-locked result buffers supplied by the caller (out=): D2H at full PCIe rate
        so = torch.empty(y.shape, dtype=torch.float32, pin_memory=True).numpy()
        xo = torch.empty(x.shape, dtype=torch.float32, pin_memory=True).numpy()
        A.project(xh_np, out=so)  # warm-up (allocates the staging buffers inside the plan)
        A.back_project(sh_np, out=xo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            A.project(xh_np, out=so)        # host in -> host out: H2D, kernels, D2H; returns when `so` is complete
            A.back_project(so, out=xo)      # its input is the forward's host result: the two calls cannot overlap
        barrier()
        dt = reduce_max((time.perf_counter() - t0) / args.steps)
        # the same two applications on INDEPENDENT host inputs, both in flight (wait=False + host_wait): the
        # adjoint's H2D runs under the forward's kernels and D2H, all three streams stay busy across the boundary
        A.project(xh_np, out=so, wait=False)
        A.back_project(sh_np, out=xo, wait=False)
        A.host_wait()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            A.project(xh_np, out=so, wait=False)
            A.back_project(sh_np, out=xo, wait=False)
            A.host_wait()
        barrier()
        dt_ov = reduce_max((time.perf_counter() - t0) / args.steps)
        per_dir = 4 * (x.numel() + y.numel())
        e2e = {"value": updates_step / dt, "unit": UNIT, "h2d_bytes_per_step": int(per_dir * world),
               "d2h_bytes_per_step": int(per_dir * world), "ms_per_step": dt * 1e3,
               "pcie_gb_per_s_per_gpu_per_direction": per_dir / dt / 1e9,
               "api": "XRayTransform3D.project/.back_project(host array, out=pinned host array) -> xct_forward_host/xct_adjoint_host; "
                      "the adjoint's input is the forward's host result",
               "overlapped": {"value": updates_step / dt_ov, "ms_per_step": dt_ov * 1e3,
                              "pcie_gb_per_s_per_gpu_per_direction": per_dir / dt_ov / 1e9,
                              "api": "project(x_host, out=, wait=False); back_project(y_host, out=, wait=False); host_wait() -> "
                                     "xct_forward_host_async / xct_adjoint_host_async / xct_host_wait (independent inputs)"},
               "numa": numa_note}

    # second half of BASELINE.json's metric: TV-regularised PDHG iterations/s on the same operator
    # (device-resident state, iteration statistics off: no host sync inside an iteration)
    solver = None
    if args.solver_iters > 0:
        from scico_b200.optimize import TVPDHG

        xh = sh = so = xo = xh_np = sh_np = None  # release the pinned e2e buffers
        S = TVPDHG(SA if world > 1 else A, y, lam=0.1, tau=0.01, sigma=0.01, maxiter=args.solver_iters)
        S.step()  # warm-up
        barrier()
        s0, s1 = ev(), ev()
        s0.record()
        for _ in range(args.solver_iters):
            S.step()
        s1.record()
        barrier()
        it_ms = reduce_max(s0.elapsed_time(s1)) / args.solver_iters
        solver = {"algorithm": "PDHG, 1/2||Ax-y||^2 + lam ||Dx||_{2,1} (scico/optimize/_primaldual.py:219-231)",
                  "iters_per_s": 1e3 / it_ms, "ms_per_iter": it_ms, "iters_timed": args.solver_iters,
                  "per_iter": "1 back projection + 1 forward projection + 3 fused TV kernels", "itstats": "off",
                  "state": "x, xbar, A^T z (volume), z1 (3 x volume), z0, y, A xbar (sinogram) resident in HBM"}
        del S
        # the same iteration with iteration statistics ON, as the reference's itstat_options compute them
        # every iteration (objective and residuals: accumulated by the iteration's own kernels, A x from A xbar by
        # linearity instead of one more forward projection, one host read of five doubles per iteration)
        if not world > 1 or hasattr(SA, "project"):
            try:
                S = TVPDHG(SA if world > 1 else A, y, lam=0.1, tau=0.01, sigma=0.01, maxiter=args.solver_iters, itstat=True)
                S.step()
                barrier()
                s0, s1 = ev(), ev()
                s0.record()
                for _ in range(args.solver_iters):
                    S.step()
                assert len(S.history) == args.solver_iters + 1  # the statistics' read-back is inside the timed region
                s1.record()
                barrier()
                on_ms = reduce_max(s0.elapsed_time(s1)) / args.solver_iters
                solver["itstats_on"] = {"iters_per_s": 1e3 / on_ms, "ms_per_iter": on_ms, "iters_timed": args.solver_iters,
                                        "per_iter": "the iteration above with objective / residual sums fused into its kernels + 1 TV-norm kernel; "
                                                    "the sums are read back once, after the timed iterations"}
                del S
            except Exception as exc:  # a statistics path that fails must not take the headline down
                solver["itstats_on"] = {"error": repr(exc)[:200]}
        # the ADMM variants of the metric's name: proximal ADMM (ct_3d_tv_padmm.py) and ADMM + CG (ct_tv_admm.py)
        from scico_b200.optimize import TVADMM, TVProximalADMM

        def time_steps(S, n):
            S.step()  # warm-up
            barrier()
            a, b = ev(), ev()
            a.record()
            for _ in range(n):
                S.step()
            if getattr(S, "itstat", False):
                assert len(S.history) == n + 1  # the statistics' read-back is inside the timed region
            b.record()
            barrier()
            return reduce_max(a.elapsed_time(b)) / n

        S = TVProximalADMM(SA if world > 1 else A, y, lam=2.0, rho=5e-3, mu=1.3e6, nu=1.01, alpha=1e2, maxiter=args.solver_iters)
        it_ms = time_steps(S, args.solver_iters)
        solver["padmm"] = {"algorithm": "ProximalADMM, A=(C; alpha D), B=-I (scico/optimize/_padmm.py:349-363, ct_3d_tv_padmm.py)",
                           "iters_per_s": 1e3 / it_ms, "ms_per_iter": it_ms, "iters_timed": args.solver_iters,
                           "per_iter": "1 back projection + 1 forward projection + 3 fused kernels", "itstats": "off"}
        del S
        try:  # with the reference's default statistics (objective, primal residual, fast dual residual) every iteration
            S = TVProximalADMM(SA if world > 1 else A, y, lam=2.0, rho=5e-3, mu=1.3e6, nu=1.01, alpha=1e2,
                               maxiter=args.solver_iters, itstat=True)
            on_ms = time_steps(S, args.solver_iters)
            solver["padmm"]["itstats_on"] = {"iters_per_s": 1e3 / on_ms, "ms_per_iter": on_ms, "iters_timed": args.solver_iters,
                                             "per_iter": "the iteration above with the statistics' sums fused into its two prox "
                                                         "kernels (fast_dual_residual, the reference's default); read back once, after the timed iterations"}
            del S
        except Exception as exc:
            solver["padmm"]["itstats_on"] = {"error": repr(exc)[:200]}
        cg_it = 2
        S = TVADMM(SA if world > 1 else A, y, lam=2.0, rho=5.0, maxiter=1, cg_tol=1e-30, cg_maxiter=cg_it)
        it_ms = time_steps(S, 1)
        solver["admm_cg"] = {"algorithm": "ADMM + CG x-step (scico/optimize/_admm.py:334-378, _admmaux.py:231-269, solver.py:367-405)",
                             "iters_per_s": 1e3 / it_ms, "ms_per_iter": it_ms, "iters_timed": 1, "cg_iters_per_admm_iter": cg_it,
                             "per_iter": f"{cg_it + 1} forward + {cg_it + 1} back projections + {3 * cg_it + 3} fused kernels; "
                                         "host reads one scalar per CG iteration (termination test)", "itstats": "off"}
        del S

    # the other BASELINE.json configurations that fit one GPU (parity for them lives in tests/; these
    # are timings only): C2 = configs[1], C3 = configs[2] on one GPU, C4 = configs[3] on one GPU
    others = None
    if world == 1 and not args.no_configs:
        x = y = None
        torch.cuda.empty_cache()
        others = other_configs(torch, sb, dev)

    # the communicating partition (view blocks), at this N
    view_block = None
    if not args.no_view_block:
        x = y = None
        torch.cuda.empty_cache()
        view_block = view_block_configs(torch, dist, sb, sharded, dev, world, barrier, reduce_max)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = cpu_sample(wl)
        sample["V_total"] = V
        v, sec, threads = time_cpu_port(sample)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample["desc"], "seconds": sec}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"XRayTransform3D {N[0]}x{N[1]}x{N[2]} volume, {V} views about axis 0, detector {D[0]}x{D[1]}, "
                                   "one forward + one adjoint per step, dense standard-normal input",
                       "partition": f"{world} z-slab(s), no collective", "kernel_path": info["path_name"],
                       "l2": "no flush: per-rank volume and sinogram (>= 0.5 GB each at 8 GPUs) exceed the 126 MB L2",
                       "fwd_ms": fwd_ms, "adj_ms": adj_ms},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "phantom": {"input": f"tanglecube phantom, {100 * nz:.1f} % non-zero voxels (the forward skips all-zero flushes)",
                        "ms_per_step": p_total / args.steps, "fwd_ms": p_fwd, "adj_ms": p_adj,
                        "value": updates_step / (p_total / args.steps * 1e-3), "unit": UNIT},
            "view_block": view_block, "solver": solver, "other_configs": others,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
