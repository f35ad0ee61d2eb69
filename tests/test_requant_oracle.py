"""Known-answer tests of the requant oracle (oracle_requant_s32_to_s8), hand-derived from the instruction
semantics of aarch64-int8/int8kernel_m4.S:386-426: scvtf (RNE), fmul, fadd (separate roundings), fcvtas
(nearest, ties AWAY from zero, saturating, NaN -> 0), sqxtn x2 (saturating narrow).  The reference holds no
vectors for these kernels (parity unpinned, see oracle.h), so the answers below are worked by hand.  CPU only."""
import numpy as np

import _libs


def rq(oracle, acc, scale, bias=None):
    c = np.array([acc], np.int32).reshape(1, -1)
    b = None if bias is None else np.array([bias], np.float32)
    return _libs.requant_s8(oracle, c, np.array([scale], np.float32), b)[0].tolist()


def test_ties_round_away_from_zero(oracle):
    # x * 0.5: 1 -> 0.5 -> 1, 3 -> 1.5 -> 2, 5 -> 2.5 -> 3 (ties-to-even would give 0, 2, 2); negatives mirror
    assert rq(oracle, [1, 3, 5, -1, -3, -5, 0, 2, -2], 0.5) == [1, 2, 3, -1, -2, -3, 0, 1, -1]


def test_below_half_is_not_rounded_up(oracle):
    # 0.49999997f is the float just below 0.5: adding 0.5 then truncating would give 1; fcvtas gives 0
    assert rq(oracle, [1, -1], np.float32(0.49999997)) == [0, 0]


def test_saturation_and_double_narrow(oracle):
    assert rq(oracle, [127, 128, 129, -128, -129, 40000, -40000, 2**31 - 1, -2**31], 1.0) == \
        [127, 127, 127, -128, -128, 127, -128, 127, -128]
    # float result beyond int32: fcvtas saturates first, then both sqxtn
    assert rq(oracle, [2**31 - 1, -2**31], 4.0) == [127, -128]


def test_bias_null_and_bias_added_after_scale(oracle):
    assert rq(oracle, [10, 11], 0.25) == [3, 3]                    # 2.5 -> 3, 2.75 -> 3
    assert rq(oracle, [10, 11], 0.25, bias=-0.5) == [2, 2]         # 2.0 -> 2, 2.25 -> 2
    assert rq(oracle, [0, 0], 123.0, bias=7.5) == [8, 8]           # K = 0 style: round(bias), tie away


def test_two_roundings_not_fused(oracle):
    # fmul then fadd, each rounded: 3 * fl(1/3) = 1.00000003 rounds to exactly 1.0f, and 1.0f - 0.5f = 0.5 is a
    # tie that goes AWAY from zero
    s = np.float32(1.0) / np.float32(3.0)
    assert np.float32(np.float32(3.0) * s) == np.float32(1.0)
    assert rq(oracle, [3], s, bias=-0.5) == [1]
    assert rq(oracle, [-3], s, bias=0.5) == [-1]


def test_nan_and_inf(oracle):
    assert rq(oracle, [5], np.float32("nan")) == [0]
    assert rq(oracle, [5, -5], np.float32("inf")) == [127, -128]
    assert rq(oracle, [0], np.float32("inf")) == [0]                 # 0 * inf = NaN -> 0


def test_scvtf_rounds_large_accumulators_to_even(oracle):
    # 2^24 + 1 is not a float: scvtf gives 2^24 (even); times 2^-17 = 128 -> saturates to 127;
    # times 2^-18 = 64 exactly
    assert rq(oracle, [2**24 + 1], np.float32(2.0 ** -18)) == [64]
    # 2^25 + 3 -> RNE to 2^25 + 4; * 2^-20 = 32.000004 -> 32
    assert rq(oracle, [2**25 + 3], np.float32(2.0 ** -20)) == [32]


def test_rowwise_scales(oracle):
    c = np.array([[100, -100, 50], [100, -100, 50]], np.int32)
    out = _libs.requant_s8(oracle, c, np.array([0.01, 1.0], np.float32), np.array([0.0, -0.25], np.float32))
    assert out.tolist() == [[1, -1, 1], [100, -100, 50]]              # 0.5 -> 1 (away); 49.75 -> 50
