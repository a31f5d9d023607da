"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/miniAero.cell, see
oracle/build_ref.sh) — full-precision conserved variables after 1, 2, 100 (and the test's own) time
steps for every case in tests/cases.py, plus the reference's 6-digit gold files re-encoded as arrays.

Run here (needs /root/reference for the gold text files and a built oracle/_ref):
    python tests/golden/make_golden.py
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

REF_TESTS = "/root/reference/kokkos/tests"


def main():
    for name, inp in cases.all_cases().items():
        steps = sorted({1, 2, 100, inp["ntimesteps"]})
        out = {}
        for kind in ("cell", "atomics"):
            for n in steps:
                if kind == "atomics" and n not in (100, inp["ntimesteps"]):
                    continue
                o = refrun.run_reference(dict(inp, ntimesteps=n), kind=kind)
                out["%s_step%d" % (kind, n)] = refrun.solution_from_dumps(o["dumps"])
        gold = os.path.join(REF_TESTS, name, "results.gold")
        if os.path.isfile(gold):
            out["results_gold"] = np.loadtxt(gold)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, sorted(out.keys()))


if __name__ == "__main__":
    main()
