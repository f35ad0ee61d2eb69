"""The 4-bit path: MXFP4 block-scaled GEMM.

PARITY UNPINNED for this row: the reference's cuda-int4 back-end is the single word "WIP"
(cuda-int4/README.md:1), so there is no reference code, test or vector to pin against.  The oracle
(oracle/oracle.c, oracle_mxf4_*) restates the published OCP MX v1.0 format; the CPU tests below pin it to
hand-worked known answers, the GPU tests check the CUDA path against it bit for bit where the arithmetic is
exactly representable (quantiser outputs; GEMMs whose every partial sum is a small dyadic rational)."""
import ctypes as C

import numpy as np
import pytest

import _libs

E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def _o(oracle):
    oracle.oracle_e2m1_encode.restype = C.c_uint8
    oracle.oracle_e2m1_encode.argtypes = [C.c_float]
    oracle.oracle_e2m1_decode.restype = C.c_float
    oracle.oracle_e2m1_decode.argtypes = [C.c_uint8]
    oracle.oracle_ue8m0_from_max.restype = C.c_uint8
    oracle.oracle_ue8m0_from_max.argtypes = [C.c_float]
    oracle.oracle_ue8m0_value.restype = C.c_double
    oracle.oracle_ue8m0_value.argtypes = [C.c_uint8]
    return oracle


def quant(oracle, x):
    """oracle quantisation of the rows of x: (q bytes [rows, kpad/2], plain scales [rows, kpad/32], kpad)."""
    x = np.ascontiguousarray(x, np.float32)
    rows, cols = x.shape
    kpad = (cols + 127) // 128 * 128
    q = np.zeros((rows, kpad // 2), np.uint8)
    sf = np.zeros((rows, kpad // 32), np.uint8)
    oracle.oracle_mxf4_quantize(rows, cols, _libs.P(x), cols, _libs.P(q), _libs.P(sf))
    return q, sf, kpad


def atoms(oracle, sf, rows, kpad):
    out = np.zeros(((rows + 127) // 128) * (kpad // 128) * 512, np.uint8)
    oracle.oracle_mxf4_sf_to_atoms(rows, kpad, _libs.P(np.ascontiguousarray(sf)), _libs.P(out))
    return out


def ref_gemm(oracle, qa, sa, qb, sb, kpad):
    m, n = qa.shape[0], qb.shape[0]
    c = np.zeros((m, n), np.float64)
    oracle.oracle_mxf4_gemm(m, n, kpad, _libs.P(qa), _libs.P(sa), _libs.P(qb), _libs.P(sb), _libs.P(c))
    return c


# ---- CPU: the oracle against hand-worked answers -----------------------------------------------------
def test_e2m1_grid_ties_and_saturation(oracle):
    o = _o(oracle)
    for code, v in enumerate(E2M1):
        assert o.oracle_e2m1_decode(code) == v and o.oracle_e2m1_decode(code | 8) == -v
        assert o.oracle_e2m1_encode(v) == code
        if v:
            assert o.oracle_e2m1_encode(-v) == (code | 8)
    # ties go to the even mantissa (0, 1, 2, 4): OCP MX v1.0 roundTiesToEven
    for x, want in [(0.25, 0.0), (0.75, 1.0), (1.25, 1.0), (1.75, 2.0), (2.5, 2.0), (3.5, 4.0), (5.0, 4.0),
                    (0.26, 0.5), (0.74, 0.5), (2.51, 3.0), (4.99, 4.0), (5.01, 6.0), (100.0, 6.0), (-7.0, -6.0)]:
        assert o.oracle_e2m1_decode(o.oracle_e2m1_encode(x)) == want, x
    assert o.oracle_e2m1_encode(float("nan")) == 0


def test_shared_scale_is_floor_log2_max_minus_two(oracle):
    o = _o(oracle)
    for mx, e in [(1.0, 125), (6.0, 127), (4.0, 127), (7.9, 127), (8.0, 128), (0.5, 124), (3.0e-5, 127 - 16 - 2)]:
        assert o.oracle_ue8m0_from_max(mx) == e, mx
    assert o.oracle_ue8m0_from_max(0.0) == 0 and o.oracle_ue8m0_from_max(float("inf")) == 255
    assert o.oracle_ue8m0_value(127) == 1.0 and o.oracle_ue8m0_value(130) == 8.0 and np.isnan(o.oracle_ue8m0_value(255))


def test_quantize_block_known_answer(oracle):
    o = _o(oracle)
    x = np.zeros((1, 40), np.float32)
    x[0, :8] = [6.0, -3.0, 1.5, 0.5, 0.2, -0.3, 12.0, 0.0]          # max 12 -> scale 2^(3-2) = 2
    x[0, 32:40] = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]          # max 8  -> scale 2, elements /2
    q, sf, kpad = quant(o, x)
    assert kpad == 128 and q.shape == (1, 64) and sf.shape == (1, 4)
    assert list(sf[0]) == [128, 128, 0, 0]
    dec = lambda b: (o.oracle_e2m1_decode(int(b) & 15), o.oracle_e2m1_decode(int(b) >> 4))
    first = [v for b in q[0, :4] for v in dec(b)]
    assert first == [3.0, -1.5, 1.0, 0.0, 0.0, -0.0, 6.0, 0.0]      # 0.75 -> 1 (tie to even), 0.25 -> 0, 0.1 -> 0, 0.15 -> 0
    second = [v for b in q[0, 16:20] for v in dec(b)]
    assert second == [0.5, 1.0, 1.5, 2.0, 2.0, 3.0, 4.0, 4.0]       # 2.5 -> 2 and 3.5 -> 4 (ties to even)
    assert not q[0, 20:].any()
    a = atoms(o, sf, 1, kpad)
    assert a[0] == 128 and a[1] == 128 and a[2] == 0 and not a[4:].any()


def test_oracle_gemm_is_exact_on_dyadic_operands(oracle):
    o = _o(oracle)
    rng = np.random.default_rng(0)
    a = rng.choice(np.array(E2M1 + [-v for v in E2M1], np.float32), (5, 256)) * 4.0    # scale 2^2 per block
    b = rng.choice(np.array(E2M1 + [-v for v in E2M1], np.float32), (7, 256)) * 0.5
    a[:, 0], b[:, 0] = 24.0, 3.0        # pin every block's maximum so the values above survive quantisation unchanged
    a[:, 32::32], b[:, 32::32] = 24.0, 3.0
    qa, sa, kpad = quant(o, a)
    qb, sb, _ = quant(o, b)
    assert np.array_equal(ref_gemm(o, qa, sa, qb, sb, kpad), a.astype(np.float64) @ b.astype(np.float64).T)


# ---- GPU: the CUDA path against the oracle -------------------------------------------------------------
gpu = pytest.mark.gpu


def _torch():
    return pytest.importorskip("torch")


@gpu
@pytest.mark.parametrize("rows,cols", [(1, 32), (128, 128), (130, 200), (300, 1000), (77, 4096)])
def test_quantizers_bit_exact(gemm, oracle, rows, cols):
    torch = _torch()
    o = _o(oracle)
    rng = np.random.default_rng(rows * 1000 + cols)
    x = (rng.standard_normal((rows, cols)) * np.exp2(rng.integers(-12, 12, (rows, 1)))).astype(np.float32)
    x[0, : min(cols, 8)] = [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 0.0][: min(cols, 8)]
    x[rows // 2] = 0.0                                               # an all-zero row: smallest scale, zero elements
    q_ref, sf_ref, kpad = quant(o, x)
    q, sf, r, k = gemm.mxf4_quantize(torch.from_numpy(x).cuda())
    assert (r, k) == (rows, cols) and gemm.last_kernel() == "mxf4_quantize_rows"
    assert np.array_equal(q.cpu().numpy().reshape(rows, kpad // 2), q_ref)
    assert np.array_equal(sf.cpu().numpy(), atoms(o, sf_ref, rows, kpad))
    # the transposing quantiser (row-major B, k x n) == the row quantiser applied to B^T
    qt, sft, r2, k2 = gemm.mxf4_quantize(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), transpose=True)
    assert (r2, k2) == (rows, cols) and gemm.last_kernel() == "mxf4_quantize_cols_t"
    assert np.array_equal(qt.cpu().numpy().reshape(rows, kpad // 2), q_ref)
    assert np.array_equal(sft.cpu().numpy(), atoms(o, sf_ref, rows, kpad))


@gpu
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (128, 128, 256), (128, 128, 1024), (256, 384, 512), (130, 70, 384),
                                   (1, 1, 32), (300, 260, 1000), (1024, 1024, 2048)])
def test_gemm_mxf4_bit_exact_on_dyadic_operands(gemm, oracle, m, n, k):
    """Every element is an E2M1 value times a per-block power of two within 2^+-1, so every product is a multiple
    of 2^-4 no larger than 144 and every partial sum (K <= 4096) a multiple of 2^-4 below 2^24 * 2^-4: exactly
    representable in fp32 whatever the summation order or the adder's alignment — the tensor-core result must equal
    the oracle's double-precision sum bit for bit."""
    torch = _torch()
    o = _o(oracle)
    rng = np.random.default_rng(m + 7 * n + 13 * k)
    grid = np.array(E2M1 + [-v for v in E2M1], np.float32)

    def operand(rows):
        x = rng.choice(grid, (rows, k))
        s = np.exp2(rng.integers(-1, 2, (rows, (k + 31) // 32))).astype(np.float32)
        x[:, ::32] = 6.0                                             # every block's maximum element: scale = s exactly
        return (x * np.repeat(s, 32, axis=1)[:, :k]).astype(np.float32)

    a, bt = operand(m), operand(n)
    qa_ref, sa_ref, kpad = quant(o, a)
    qb_ref, sb_ref, _ = quant(o, bt)
    want = ref_gemm(o, qa_ref, sa_ref, qb_ref, sb_ref, kpad)
    assert np.array_equal(want, a.astype(np.float64) @ bt.astype(np.float64).T)      # the fixture loses nothing to quantisation
    qa, sfa, _, _ = gemm.mxf4_quantize(torch.from_numpy(a).cuda())
    qb, sfb, _, _ = gemm.mxf4_quantize(torch.from_numpy(np.ascontiguousarray(bt.T)).cuda(), transpose=True)
    C_ = torch.full((m, n), float("nan"), device="cuda")
    gemm.gemm_mxf4(qa, sfa, qb, sfb, m, n, k, out=C_)
    assert gemm.last_kernel() == "tc_mxf4_128x128"
    got = C_.cpu().numpy()
    assert np.array_equal(got.astype(np.float64), want), np.abs(got - want).max()


@gpu
@pytest.mark.parametrize("m,n,k", [(200, 136, 264), (512, 768, 4096)])
def test_gemm_mxf4_random_operands(gemm, oracle, m, n, k):
    """uniform(-1,1) operands (cuda/random_matrix.cpp) quantised on the device: the GEMM of the quantised operands
    against the oracle's double sum (fp32 accumulation noise only), and the quantisation error itself for the record."""
    torch = _torch()
    o = _o(oracle)
    a, b = _libs.gen_f32(o, m, k, 81), _libs.gen_f32(o, k, n, 82)
    qa, sfa, _, _ = gemm.mxf4_quantize(torch.from_numpy(a).cuda())
    qb, sfb, _, _ = gemm.mxf4_quantize(torch.from_numpy(b).cuda(), transpose=True)
    got = gemm.gemm_mxf4(qa, sfa, qb, sfb, m, n, k).cpu().numpy()
    qa_ref, sa_ref, kpad = quant(o, a)
    qb_ref, sb_ref, _ = quant(o, np.ascontiguousarray(b.T))
    want = ref_gemm(o, qa_ref, sa_ref, qb_ref, sb_ref, kpad)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max() * max(1.0, (k / 1024) ** 0.5)
    full = _libs.ref_f64(o, a, b)
    assert np.abs(got - full).max() <= 0.25 * np.abs(full).max()     # 4-bit operands: a coarse GEMM by construction


@gpu
def test_gemm_mxf4_arguments(gemm):
    torch = _torch()
    z = torch.zeros(64, dtype=torch.uint8, device="cuda")
    Cm = torch.full((4, 4), 3.0, device="cuda")
    assert gemm.lib.b200_gemm_mxf4(4, 4, 0, None, None, None, None, Cm.data_ptr(), 4, None) == 0 and (Cm == 0).all()
    assert gemm.lib.b200_gemm_mxf4(4, 4, 32, None, z.data_ptr(), z.data_ptr(), z.data_ptr(), Cm.data_ptr(), 4, None) == -1
    assert gemm.lib.b200_gemm_mxf4(4, 4, 32, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), Cm.data_ptr(), 3, None) == -1
    assert gemm.lib.b200_mxf4_q_bytes(3, 100) == 3 * 64 and gemm.lib.b200_mxf4_sf_bytes(129, 100) == 2 * 512
