"""Run one case of tests/test_gpu_ipc_one_gpu.py with every rank's traceback written to gpurun_out/ipc_rank<r>.log
(mp.spawn reports only the first rank that fails, which is often a rank that lost its peer, not the culprit).
Usage: python tools/debug_ipc_case.py <case> <world> [exchange] [rendezvous]"""
import os
import sys
import traceback

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def _cases():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ipc_cases", os.path.join(ROOT, "tests", "test_gpu_ipc_one_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(rank, world, port, case, exchange, rendezvous):
    _worker = _cases()._worker
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/ipc_rank{rank}.log", "w") as fh:
        try:
            _worker(rank, world, port, case, exchange, rendezvous)
            fh.write("ok\n")
        except BaseException:
            fh.write(traceback.format_exc())
            raise


if __name__ == "__main__":
    import torch.multiprocessing as mp

    _free_port = _cases()._free_port
    case, world = sys.argv[1], int(sys.argv[2])
    exchange = sys.argv[3] if len(sys.argv) > 3 else "peer"
    rendezvous = sys.argv[4] if len(sys.argv) > 4 else "flags"
    mp.spawn(run, args=(world, _free_port(), case, exchange, rendezvous), nprocs=world, join=True)
