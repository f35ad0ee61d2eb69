"""The oracle against the committed golden vectors (tests/golden/reference_vectors.npz, generated
from the reference's own compiled sources by tests/golden/make_golden.py).  CPU only; runs on the
GPU box too, where /root/reference does not exist."""
import os

import numpy as np
import pytest

import _libs
from _libs import P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
F32 = sorted({k.split("_")[1] for k in G.files if k.startswith("f32_")})
S8 = sorted({k.split("_")[1] for k in G.files if k.startswith("s8_")})


@pytest.mark.parametrize("idx", F32)
def test_f32_golden(oracle, idx):
    m, n, k, seed = (int(x) for x in G[f"f32_{idx}_shape"])
    a, b = G[f"f32_{idx}_a"], G[f"f32_{idx}_b"]
    # generator: same stream as cuda/random_matrix.cpp after srand48(seed), A then B
    oracle.oracle_seed(seed)
    a2 = np.zeros(m * k, np.float32)
    b2 = np.zeros(k * n, np.float32)
    oracle.oracle_random_matrix_cuda(m, k, P(a2), m)
    oracle.oracle_random_matrix_cuda(k, n, P(b2), k)
    assert np.array_equal(a2.reshape(m, k), a) and np.array_equal(b2.reshape(k, n), b)
    # contraction: bit-exact vs the reference's naive REF_MMult, noise-level vs its OpenBLAS one
    c = _libs.ref_f32_fma(oracle, a, b)
    assert np.array_equal(c, G[f"f32_{idx}_c_naive"])
    t = _libs.ref_f64(oracle, a, b)
    assert np.abs(G[f"f32_{idx}_c_openblas"] - t).max() <= 4e-7 * k ** 0.5 * np.abs(t).max() + 1e-6
    # the max|diff| the reference harness would print (cuda/test_MMult.cpp:123) is far below its 0.5 gate
    d = oracle.oracle_compare_matrices_f32(m, n, P(c), n, P(np.ascontiguousarray(G[f"f32_{idx}_c_openblas"])), n)
    assert d < 1e-4


def test_ones_golden(oracle):
    a, b, c = G["ones_a"], G["ones_b"], G["ones_c"]
    assert (c == a.shape[1]).all()                      # every element equals K exactly
    assert np.array_equal(_libs.ref_f32_fma(oracle, a, b), c)


@pytest.mark.parametrize("idx", S8)
def test_s8_golden(oracle, idx):
    a, b, c = G[f"s8_{idx}_a"], G[f"s8_{idx}_b"], G[f"s8_{idx}_c"]
    a2 = np.zeros_like(a)
    oracle.oracle_random_int8_ramp(a.shape[0], a.shape[1], P(a2), a.shape[1])
    assert np.array_equal(a, a2)
    assert np.array_equal(_libs.ref_s8(oracle, a, b), c)


def test_bf16_rounding(oracle):
    x = np.array([1.0, 1.0 + 2 ** -8, 1.0 + 2 ** -8 + 2 ** -20, 1.0 + 3 * 2 ** -8, -0.3, 3.0e38, 1e-40], np.float32)
    got = _libs.round_bf16(oracle, x)
    import torch
    want = torch.from_numpy(x).bfloat16().float().numpy()
    assert np.array_equal(got, want)
