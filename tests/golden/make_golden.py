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


def main_parallel():
    """Per-rank full-precision results of the reference's WITH_MPI build (oracle/_ref/miniAero.cell.mpi over the
    file-based MPI stand-in) -> tests/golden/par_<case>_<N>.npz; plus the reference's parallel gold files."""
    for name, (inp, rank_counts, ref_dir, _, _) in cases.PARALLEL.items():
        for nranks in rank_counts:
            out = {"nranks": np.array(nranks)}
            for n in sorted({1, 2, 100, inp["ntimesteps"]}):
                per_rank = refrun.run_reference_parallel(dict(inp, ntimesteps=n), nranks)
                for r, o in enumerate(per_rank):
                    nown = o["results"].shape[0]
                    out["r%d_step%d" % (r, n)] = refrun.solution_from_dumps(o["dumps"])[:nown]
            for r in range(nranks):
                gold = os.path.join(REF_TESTS, ref_dir or "-", "results.%d.gold" % r)
                if ref_dir and os.path.isfile(gold) and len(os.listdir(os.path.join(REF_TESTS, ref_dir))) == nranks + 3:
                    out["r%d_gold" % r] = np.loadtxt(gold)
            np.savez_compressed(os.path.join(HERE, "par_%s_%d.npz" % (name, nranks)), **out)
            print(name, nranks, len(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "parallel":
        main_parallel()
    else:
        main()
