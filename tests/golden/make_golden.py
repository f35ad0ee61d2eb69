"""Generates tests/golden/*.npz from the REFERENCE ITSELF: oracle/_ref/libref.so is compiled from
the reference's own sources (oracle/Makefile `ref`), so every array below is an output of reference
code run in the build container.  /root/reference does not exist on the GPU box; these fixtures do.

    python tests/golden/make_golden.py

fp32 cases:  inputs from cuda/random_matrix.cpp after srand48(seed) (called as the harness calls it,
             cuda/test_MMult.cpp:77-78), outputs of cuda/REF_MMult.cpp (OpenBLAS cblas_sgemm) and of
             aarch64/REF_MMult.cpp (naive, fused by the reference's own flags).
ones case:   aarch64/random_matrix.cpp (all 1.0f) -> every C element == K (SURVEY §8c fixture 3).
int8 cases:  aarch64-int8/random_matrix.c ramp + aarch64-int8/REF_MMult.c.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402

P = _libs.P


def main():
    r = _libs.load_ref()
    r.openblas_set_num_threads(1)
    libc = C.CDLL(None)
    libc.srand48.argtypes = [C.c_long]
    out = {}
    for idx, (m, n, k, seed) in enumerate([(64, 48, 80, 1), (96, 128, 160, 2), (130, 70, 257, 3), (128, 256, 64, 4)]):
        libc.srand48(seed)
        a = np.zeros(m * k, np.float32)
        b = np.zeros(k * n, np.float32)
        r.cuda_random_matrix(m, k, P(a), m)      # cuda/test_MMult.cpp:77
        r.cuda_random_matrix(k, n, P(b), k)      # cuda/test_MMult.cpp:78
        a, b = a.reshape(m, k), b.reshape(k, n)
        c_blas = np.zeros((m, n), np.float32)
        r.cuda_REF_MMult(m, n, k, P(a), k, P(b), n, P(c_blas), n)
        c_naive = np.zeros((m, n), np.float32)
        r.a64_REF_MMult(m, n, k, P(a), P(b), P(c_naive))
        out[f"f32_{idx}_shape"] = np.array([m, n, k, seed])
        out[f"f32_{idx}_a"], out[f"f32_{idx}_b"] = a, b
        out[f"f32_{idx}_c_openblas"], out[f"f32_{idx}_c_naive"] = c_blas, c_naive
    m = n = k = 96
    a = np.zeros((m, k), np.float32)
    b = np.zeros((k, n), np.float32)
    r.a64_random_matrix(m, k, P(a))
    r.a64_random_matrix(k, n, P(b))
    c = np.zeros((m, n), np.float32)
    r.a64_REF_MMult(m, n, k, P(a), P(b), P(c))
    out["ones_a"], out["ones_b"], out["ones_c"] = a, b, c
    for idx, (m, n, k) in enumerate([(77, 77, 77), (64, 96, 128), (5, 130, 33)]):
        a = np.zeros((m, k), np.int8)
        b = np.zeros((k, n), np.int8)
        r.i8_random_matrix(m, k, P(a), k)
        r.i8_random_matrix(k, n, P(b), n)
        c = np.zeros((m, n), np.int32)
        r.i8_REF_MMult(m, n, k, P(a), k, P(b), n, P(c), n)
        out[f"s8_{idx}_a"], out[f"s8_{idx}_b"], out[f"s8_{idx}_c"] = a, b, c
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
