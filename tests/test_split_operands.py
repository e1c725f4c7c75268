"""CPU check of the EXACT_TC error model (oracle/split_operands.py): the three-term split-fp16 product is as good as an
fp32 GEMM when the weight planes are scaled into the normal fp16 range, and measurably worse when they are not (conv1's
BN-folded weights are ~1e-3: their `lo` plane would be subnormal)."""
import torch

from oracle import split_operands as S


def rel(a, ref):
    return float((a.double() - ref).norm() / ref.norm())


def gemm(seed, M, K, N, w_scale):
    g = torch.Generator().manual_seed(seed)
    a = torch.relu(torch.randn(M, K, generator=g)) * 3.0          # post-ReLU activations
    b = torch.randn(K, N, generator=g) * w_scale
    return a, b, a.double() @ b.double()


def test_split_product_matches_fp32_when_weights_are_scaled():
    for seed, w_scale in ((1, 1.0e-3), (2, 5.0e-2), (3, 2.0)):
        a, b, ref = gemm(seed, 256, 576, 96, w_scale)
        e32 = rel(a @ b, ref)
        e_split = rel(S.split_matmul(a, b), ref)
        e16 = rel(S.fp16_matmul(a, b), ref)
        assert e_split < 1.0e-6 and e_split < 4.0 * e32 + 1.0e-7, (w_scale, e_split, e32)
        assert e16 > 50.0 * e_split, (w_scale, e16, e_split)       # plain fp16 operands: ~3e-4


def test_unscaled_small_weights_lose_the_lo_plane():
    a, b, ref = gemm(4, 256, 147, 64, 1.0e-3)                       # conv1-like: 7x7x3 taps, folded weights ~1e-3
    e_scaled = rel(S.split_matmul(a, b, scale_weights=True), ref)
    e_unscaled = rel(S.split_matmul(a, b, scale_weights=False), ref)
    assert e_scaled < 1.0e-6
    assert e_unscaled > 10.0 * e_scaled, (e_unscaled, e_scaled)


def test_weight_scale_is_a_power_of_two_into_the_target_range():
    for m in (1.0e-4, 3.3e-3, 0.7, 1.0, 5000.0, 8192.0):
        s = S.weight_scale(torch.tensor([m, -m / 3]))
        assert 4096.0 <= m * s < 8192.0
        mant, _ = __import__("math").frexp(s)
        assert mant == 0.5                                          # exact power of two: scaling loses no bits
    assert S.weight_scale(torch.zeros(3)) == 1.0


def test_dropped_lo_lo_term_is_below_fp32_rounding():
    a, b, ref = gemm(5, 128, 1152, 128, 2.0e-2)
    s = S.weight_scale(b)
    a_hi, a_lo = S.split(a)
    b_hi, b_lo = S.split(b * s)
    three = (a_lo.double() @ b_hi.double() + a_hi.double() @ b_lo.double() + a_hi.double() @ b_hi.double()) / s
    four = three + (a_lo.double() @ b_lo.double()) / s
    assert rel(three.float(), ref) < 5.0e-7
    assert abs(rel(four.float(), ref) - rel(three.float(), ref)) < 5.0e-8
