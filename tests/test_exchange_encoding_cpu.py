"""The number formats the tm2 recurrent kernels (csrc/lstm_recurrent_tm2.cu) put on the wire and into the tensor cores, restated in
numpy and checked on the CPU: the fp16 two-term split (t2_split), the forward exchange word with its tag bit, the tagged BPTT partial.
These are the invariants the kernels assume; the GPU parity tests check the kernels themselves."""
import numpy as np

LO_SCALE = np.float32(2048.0)


def split16(x):
    """t2_split: hi = fp16(x), lo' = fp16((x - hi) * 2^11), both round-to-nearest-even; x - hi and the scaling are exact in fp32."""
    x = np.asarray(x, np.float32)
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * LO_SCALE).astype(np.float16)
    return hi, lo


def test_two_term_split_carries_22_bits():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 200000), rng.standard_normal(200000) * 0.1, rng.uniform(-1, 1, 200000) * 1e-3,
                        np.array([0.0, 1.0, -1.0, 0.5, 6.1e-5, 2.0 ** -14, 2.0 ** -24, 65000.0])]).astype(np.float32)
    hi, lo = split16(x)
    back = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
    err = np.abs(back - x.astype(np.float64))
    # relative 2^-22 for values fp16 represents as normal numbers, an absolute floor of 2^-36 below
    assert np.all(err <= np.maximum(2.0 ** -22 * np.abs(x), 2.0 ** -36))
    # fp16 and tf32 have the same 11 significant bits: the dropped lo*lo term is the only second-order loss
    assert np.all(np.abs(lo.astype(np.float32)) <= np.maximum(np.abs(x), 2.0 ** -13))


def test_forward_exchange_word_has_a_free_tag_bit():
    """|h| <= 1 (tanh * logistic) gives |lo'| <= 1/2: bit 14 of the lo' half (bit 30 of the word) is never set by data, so the step
    tag rides there and `& 0xBFFF` takes it out again; hi and lo' survive the round trip bit for bit."""
    rng = np.random.default_rng(1)
    h = np.concatenate([rng.uniform(-1, 1, 500000), np.array([1.0, -1.0, 0.0, 2.0 ** -25, 1 - 2.0 ** -24])]).astype(np.float32)
    hi, lo = split16(h)
    lo_bits = lo.view(np.uint16)
    assert not np.any(lo_bits & 0x4000)
    for tag in (0, 1):
        word = hi.view(np.uint16).astype(np.uint32) | (((lo_bits & 0xBFFF).astype(np.uint32) | (tag << 14)) << 16)
        assert np.all(((word >> 30) & 1) == tag)
        assert np.array_equal((word & 0xFFFF).astype(np.uint16), hi.view(np.uint16))
        assert np.array_equal(((word >> 16) & 0xBFFF).astype(np.uint16), lo_bits)
    # cudaMemset(0x40) arms every word with tag 1, the first two steps of a pass carry tag 0
    assert (0x40404040 >> 30) & 1 == 1


def tag_partial(v, tag):
    """BPTT exchange: last mantissa bit rounded away to nearest even (inf / NaN keep their class), then replaced by the tag."""
    b = np.asarray(v, np.float32).view(np.uint32).astype(np.uint64)
    special = (b & 0x7F800000) == 0x7F800000
    r = np.where(special, b & ~np.uint64(1), (b + ((b >> np.uint64(1)) & np.uint64(1))) & ~np.uint64(1))
    return (r | np.uint64(tag)).astype(np.uint32)


def test_tagged_partial_is_an_unbiased_one_ulp_rounding():
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.standard_normal(400000) * 10.0 ** rng.integers(-12, 4, 400000),
                        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 3.4028235e38, 1.17549435e-38, 1e-45])]).astype(np.float32)
    for tag in (0, 1):
        w = tag_partial(v, tag)
        assert np.all((w & 1) == tag)
        back = (w & ~np.uint32(1)).view(np.float32)
        fin = np.isfinite(v)
        ulp = np.spacing(np.abs(v[fin]))
        assert np.all(np.abs(back[fin].astype(np.float64) - v[fin].astype(np.float64)) <= ulp.astype(np.float64) * 1.0000001)
        assert np.all(np.isnan(back[np.isnan(v)])) and np.array_equal(np.isinf(back), np.isinf(v) | (np.abs(v) >= 3.4028234e38))
    # round to nearest EVEN: the dropped bit does not bias the sum the way truncation would
    w = (tag_partial(v[:400000], 0) & ~np.uint32(1)).view(np.float32)
    rel = (w.astype(np.float64) - v[:400000].astype(np.float64)) / np.abs(v[:400000].astype(np.float64))
    assert abs(rel.mean()) < 2.0 ** -24 * 0.02
    # cudaMemset(0x01) arms every word with tag 1
    assert 0x01010101 & 1 == 1


def test_delta_operand_scale_keeps_small_deltas_normal():
    """BPTT B operand: deltas (|d| <= 1 after limitedError) are scaled by 2^13 before the split, so deltas down to 7e-9 keep a normal
    fp16 hi half; unscaled, fp16's 5-bit exponent would put an absolute floor of 1.5e-11 under them (1.5e-4 relative at 1e-7)."""
    d = np.float32(1e-7) * np.random.default_rng(3).uniform(0.5, 1.0, 100000).astype(np.float32)
    hi, lo = split16(d * np.float32(8192.0))
    back = (hi.astype(np.float64) + lo.astype(np.float64) / 2048.0) / 8192.0
    assert np.max(np.abs(back - d.astype(np.float64)) / d) <= 2.0 ** -21
    hi0, lo0 = split16(d)
    back0 = hi0.astype(np.float64) + lo0.astype(np.float64) / 2048.0
    assert np.max(np.abs(back0 - d.astype(np.float64)) / d) > 1e-5        # what the scale avoids
    assert np.max(np.abs(hi.astype(np.float32))) < 65504 and 8192.0 * 1.0 < 65504


def test_two_term_step_product_reaches_fp32_accuracy():
    """W h ~= W_hi h_hi + 2^-11 (W_hi h_lo' + W_lo' h_hi), products of fp16 pairs (exact in fp32) summed in fp32: the scheme of the step GEMM
    (32 kind::f16 MMAs per 256-wide step), emulated here; tools/micro/tcgen05_f16_step.cu measured the same on the tensor cores
    (profiles/r02_probes.txt: 3.9e-7 .. 5.1e-7 of max|W h|, the reference's own serial fp32 sum 3.8e-7 .. 5.0e-7)."""
    rng = np.random.default_rng(5)
    for scale in (1e-4, 0.1, 3.0):
        W = rng.uniform(-scale, scale, (128, 256)).astype(np.float32)
        h = (np.tanh(rng.standard_normal((256, 16))) * rng.uniform(0, 1, (256, 16))).astype(np.float32)
        ref = W.astype(np.float64) @ h.astype(np.float64)
        # per-CTA power-of-two weight scale (t2_slice_scale) is only needed beyond fp16 range; these weights are inside it
        Whi, Wlo = split16(W)
        hhi, hlo = split16(h)
        f = lambda a: a.astype(np.float32)                                            # noqa: E731
        d0 = f(Whi) @ f(hhi)
        d1 = f(Whi) @ f(hlo) + f(Wlo) @ f(hhi)
        full = d0 + d1 * np.float32(1.0 / 2048.0)
        denom = np.abs(ref).max()
        assert np.abs(full - ref).max() / denom < 1.5e-6
        assert np.abs(d0 - ref).max() / denom > 5e-5                                   # one term alone is an 11-bit product
        serial = np.zeros((128, 16), np.float32)
        for k in range(256):                                                          # the reference's own order: one fp32 sum over k
            serial += W[:, k:k + 1] * h[k:k + 1, :]
        assert np.abs(full - ref).max() <= 4 * np.abs(serial - ref).max()
