"""Generates tests/golden/vanalbada_*.npz: full-precision results of the reference with its stencil limiter switched to
VanAlbadaLimiter (oracle/_ref/miniAero.cell.vanalbada: the unmodified sources, the limiter class redirected by
oracle/vanalbada_swap.h) after 1, 2 and 100 steps, for the second-order cases of tests/cases.py.

Run here (needs /root/reference and a built oracle/_ref):  python tests/golden/make_golden_vanalbada.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import refrun  # noqa: E402

CASES = ("sod_o2", "sod_o2_visc", "ramp_odd")


def main():
    for name in CASES:
        inp = cases.EXTRA[name]
        out = {}
        for n in (1, 2, 100):
            o = refrun.run_reference(dict(inp, ntimesteps=n), kind="cell.vanalbada")
            out["cell_step%d" % n] = refrun.solution_from_dumps(o["dumps"])
        np.savez_compressed(os.path.join(HERE, "vanalbada_%s.npz" % name), **out)
        print(name, sorted(out.keys()))


if __name__ == "__main__":
    main()
