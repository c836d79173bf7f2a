"""Writes the committed fixtures of tests/golden/.

 - kat.json            known-answer vectors TRANSCRIBED from the reference (README.md:58-142,
                       src/Filters.jl:276-280, notebook cell 10); no code is run to make them.
 - oracle_vectors.npz  seeded inputs and the outputs of oracle/multirate_oracle.py for every kernel
                       type x dtype, in 3 chunks (so GPU parity tests also run against frozen data).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys
from fractions import Fraction

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import multirate_oracle as mo  # noqa: E402

kat = {
    "readme_3_17": {   # README.md:58-134
        "h": [1, 1, 1, 0, 0, 0, 0, 0, 0], "ratio": [3, 17], "x": list(range(1, 101)), "chunks": [5, 18, 77],
        "y": [[1.0], [6.0, 12.0, 18.0, 23.0], [29.0, 35.0, 40.0, 46.0, 52.0, 57.0, 63.0, 69.0, 74.0, 80.0, 86.0, 91.0, 97.0]],
        "pfb": [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]],          # README.md:88-90
        "Nphi": 3, "tapsPerphi": 3, "criticalYidx": 0, "phiIdx": 1, "inputDeficit": 1, "history": [0.0, 0.0], "historyLen": 2,
    },
    "taps2pfb_example": {"h": list(range(1, 10)), "Nphi": 4, "pfb": [[9, 0, 0, 0], [5, 6, 7, 8], [1, 2, 3, 4]]},   # src/Filters.jl:276-280
    "farrow_notebook_count": {"Nphi": 32, "tapsPerphi": 10, "polyorder": 4, "rate": float(np.pi), "n_in": 40, "n_out": 126},
    "readme_benchmark_count": {"ratio": [147, 160], "n_in": 1000000, "n_out": 918750, "bytes": 7350144},            # README.md:190-193
}
json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

rng = np.random.default_rng(20141117)
out = {}
cases = [("standard", Fraction(1, 1), None, None), ("decimator", Fraction(1, 8), None, None),
         ("interpolator", Fraction(4, 1), None, None), ("rational", Fraction(147, 160), None, None),
         ("rational_3_17", Fraction(3, 17), None, None), ("arbitrary", 0.918734, 32, None), ("farrow", 0.918734, 32, 4)]
for name, ratio, nphi, po in cases:
    for th in (np.float32, np.float64):
        for tx in (np.float32, np.complex64, np.float64, np.complex128):
            hlen = 320 if isinstance(ratio, float) else (441 if name == "rational" else 61)
            h = (rng.random(hlen) - 0.3).astype(th)
            x = rng.random((2, 331))
            if np.dtype(tx).kind == "c":
                x = x + 1j * rng.random((2, 331))
            x = x.astype(tx)
            f = mo.FIRFilter(h, ratio, nphi, po)
            ys = [f.filt(x[:, a:b]) for a, b in ((0, 1), (1, 40), (40, 331))]
            key = "%s.%s.%s" % (name, np.dtype(th).name, np.dtype(tx).name)
            out[key + ".h"] = h
            out[key + ".x"] = x
            for i, y in enumerate(ys):
                out[key + ".y%d" % i] = y
            out[key + ".state"] = np.array([f.state().get("phiIdx", 1), f.state().get("inputDeficit", 1)], dtype=np.int64)
            out[key + ".fstate"] = np.array([f.state().get("acc", 1.0), f.state().get("alpha", 0.0)], dtype=np.float64)
            if name == "farrow":
                # the Farrow coefficients are an input of the path: frozen with the vectors (the fit is solver dependent)
                out[key + ".pnfb"] = np.asarray(f.kernel.pnfb, dtype=np.float64)
np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
print("wrote", len(out), "arrays")
