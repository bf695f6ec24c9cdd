"""Golden vectors from the reference's OWN published run of its 3D CT example.

`data/notebooks/ct_3d_tv_padmm.ipynb` (the executed notebook of `examples/scripts/ct_3d_tv_padmm.py`, shipped with
the reference) holds the iteration statistics its ProximalADMM printed on an RTX 2080 Ti with real JAX / XLA: Objective,
Prml Rsdl, Dual Rsdl of all 1000 iterations (4 significant digits) and the final SNR / MAE.  This script copies those
NUMBERS (no code) into `tests/golden/nb_ct_3d_tv_padmm.npz`; `tests/test_reference_notebook.py` (oracle, CPU) and
`tests/test_gpu_reference_notebook.py` (CUDA path) rebuild the example and compare.

    python tests/golden/make_notebook_golden.py [/root/reference]
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(ref="/root/reference"):
    nb = json.load(open(os.path.join(ref, "data", "notebooks", "ct_3d_tv_padmm.ipynb")))
    txt = ""
    for cell in nb["cells"]:
        if cell["cell_type"] == "code":
            for out in cell.get("outputs", []):
                if "text" in out:
                    txt += "".join(out["text"])
    num = r"(\d\.\d+e[+-]\d+)"
    rows = re.findall(r"\s*(\d+)\s+" + r"\s+".join([num] * 4), txt)
    it = np.array([int(r[0]) for r in rows])
    assert len(rows) == 1000 and np.array_equal(it, np.arange(1000))
    m = re.search(r"SNR: ([\d.]+) \(dB\), MAE: ([\d.]+)", txt)
    np.savez_compressed(
        os.path.join(HERE, "nb_ct_3d_tv_padmm.npz"),
        time=np.array([float(r[1]) for r in rows]), objective=np.array([float(r[2]) for r in rows]),
        prml_rsdl=np.array([float(r[3]) for r in rows]), dual_rsdl=np.array([float(r[4]) for r in rows]),
        snr_db=float(m.group(1)), mae=float(m.group(2)),
        source="data/notebooks/ct_3d_tv_padmm.ipynb cell 7 (examples/scripts/ct_3d_tv_padmm.py:42-133), RTX 2080 Ti")
    print(len(rows), "iterations; SNR", m.group(1), "MAE", m.group(2))


if __name__ == "__main__":
    main(*sys.argv[1:])
