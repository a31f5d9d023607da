"""Generates tests/golden/initial_conditions.npz: the reference's state after ZERO time steps (initialize_sod3d /
initialize_constant, Initial_Conditions.h:38-133, as copied out at TimeSolverExplicitRK4.h:516) for three cases.
Run here (needs a built oracle/_ref):  python tests/golden/make_golden_ic.py
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

CASES = ("sod_o2_visc", "flatplate_o1", "ramp_odd")


def main():
    out = {}
    for name in CASES:
        o = refrun.run_reference(dict(cases.EXTRA[name], ntimesteps=0), kind="cell")
        out[name] = refrun.solution_from_dumps(o["dumps"])
        print(name, out[name].shape)
    np.savez_compressed(os.path.join(HERE, "initial_conditions.npz"), **out)


if __name__ == "__main__":
    main()
