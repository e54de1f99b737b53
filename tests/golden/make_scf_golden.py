"""Generates tests/golden/scf_fortran.npz from the reference's own SCF golden vectors
(/root/reference/tests/potential/scf/data: positions.dat.gz, <set>.coeff, <set>-accp.dat.gz,
produced by the Fortran SCF code via adrn/biff, data/README.md).  Only DATA is transcribed: the first
256 positions of each of the five coefficient sets, with the Hernquist&Ostriker -> Lowing coefficient
conversion of tests/potential/scf/test_accp_fortran.py:50-56 already applied.  Run in the build
container (the reference tree does not travel to the GPU box):  python tests/golden/make_scf_golden.py
"""
import os
from math import factorial

import numpy as np

REF = "/root/reference/tests/potential/scf/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scf_fortran.npz")
SETS = ["simple-hernquist", "multi-hernquist", "simple-nonsph", "random", "wang-zhao"]
NPOS = 256

xyz = np.loadtxt(os.path.join(REF, "positions.dat.gz"), skiprows=1)[:NPOS]
out = {"xyz": xyz}
for name in SETS:
    coeff = np.atleast_2d(np.loadtxt(os.path.join(REF, f"{name}.coeff"), skiprows=1))
    accp = np.loadtxt(os.path.join(REF, f"{name}-accp.dat.gz"))[:NPOS]
    nmax = coeff[:, 0].astype(int).max()
    lmax = coeff[:, 1].astype(int).max()
    S = np.zeros((nmax + 1, lmax + 1, lmax + 1)); T = np.zeros_like(S)
    for n, l, m, cc, sc in coeff:
        if l != 0:
            fac = np.sqrt(4 * np.pi) * np.sqrt((2 * l + 1) / (4 * np.pi) * factorial(int(l - m)) / factorial(int(l + m)))
            cc /= fac; sc /= fac
        S[int(n), int(l), int(m)] = cc; T[int(n), int(l), int(m)] = sc
    key = name.replace("-", "_")
    out[key + "_S"] = S; out[key + "_T"] = T
    out[key + "_grad"] = accp[:, :3]          # compared to gradient() with G=M=r_s=1 (test_accp_fortran.py:150-156)
    out[key + "_pot"] = -accp[:, -1]          # compared to potential() (:104-106)
np.savez_compressed(OUT, **out)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
