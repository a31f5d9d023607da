"""Generates tests/golden/unit_functions.npz: seeded inputs (tests/unit_inputs.py) and the outputs of the
reference's own device functions on them (oracle/_ref/unit_oracle = oracle/unit_oracle.cpp around the reference
headers Roe_Flux.h, Viscous_Flux.h, GasModel.h, VenkatLimiter.h, VanAlbadaLimiter.h).

Run here (needs /root/reference and a built oracle/_ref):  python tests/golden/make_golden_unit.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refrun  # noqa: E402
import unit_inputs  # noqa: E402


def main():
    out = {}
    for fn, rows in unit_inputs.make_inputs().items():
        out[fn + "_in"] = rows
        out[fn + "_out"] = refrun.unit_oracle(fn, rows)
        print(fn, rows.shape, "->", out[fn + "_out"].shape, "finite:", bool(np.isfinite(out[fn + "_out"]).all()))
    np.savez_compressed(os.path.join(HERE, "unit_functions.npz"), **out)


if __name__ == "__main__":
    main()
