"""Generates the committed golden fixtures from the reference's own ground truth (CPU LAPACK ?sygvd/?hegvd,
test_driver/test_zhegvdx.F90:163-182) on seeded inputs.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lapack, matgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for cplx in (False, True):
        for fam, gen in (("c", matgen.family_c), ("r", matgen.family_r)):
            n = 48
            a, b = gen(n, cplx, seed=2024)
            w, z, u, info = lapack.hegvd(a, b)
            assert info == 0
            np.savez_compressed(os.path.join(HERE, f"hegvd_{'z' if cplx else 'd'}_{fam}_n{n}.npz"),
                                a=a, b=b, w=w, absz=np.abs(z), u=np.triu(u))
    for kind in ("random", "wilkinson", "glued", "toeplitz"):
        n = 97
        d, e = matgen.tridiag_family(n, kind, seed=7)
        w, z, info = lapack.stedc(d, e)
        assert info == 0
        np.savez_compressed(os.path.join(HERE, f"stedc_{kind}_n{n}.npz"), d=d, e=e, w=w, absz=np.abs(z))
    for cplx in (False, True):
        n = 40
        a, _ = matgen.family_c(n, cplx, seed=11)
        c, d, e, tau = lapack.hetrd(a)
        np.savez_compressed(os.path.join(HERE, f"hetrd_{'z' if cplx else 'd'}_n{n}.npz"), a=a, d=d, e=e, tau=tau)


if __name__ == "__main__":
    main()
