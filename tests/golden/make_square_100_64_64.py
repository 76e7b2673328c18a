"""Regenerates the reference fixture `tests/square_100_64_64.mats`, which BASELINE.json's
config 2 names but which is missing from the reference mount (.MISSING_LARGE_BLOBS:6).

Recipe: 100 general matrices, entries U(0,1) like the shipped
`square_5_64_64.mats`, generated in fp64 with numpy.random.default_rng(20260101) and rounded
to fp32.  Writes the compact .npz the tests use; `--mats PATH` additionally writes the text
`.mats` form (15 significant digits, like the reference's square_* files) for the CLIs.

    python tests/golden/make_square_100_64_64.py [--mats square_100_64_64.mats]
"""
import argparse
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def generate():
    rng = np.random.default_rng(20260101)
    return rng.random((100, 64, 64)).astype(np.float32)      # a[k, i, j]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mats", default=None)
    args = ap.parse_args()
    a = generate()
    np.savez_compressed(os.path.join(HERE, "square_100_64_64.npz"), a=a)
    if args.mats:
        with open(args.mats, "w") as f:
            f.write("100 64 64\n")
            for k in range(100):
                for i in range(64):
                    f.write("  ".join(f"{v:.15g}" for v in a[k, i]) + "\n")


if __name__ == "__main__":
    main()
