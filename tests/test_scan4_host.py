"""CPU: the two bit tricks of the bulk-copy fed map scan (pytorchocr_b200/csrc/db_scan4.cuh), restated in numpy.

1. mask bit = sign bit of `thresh - f` (one FADD, shifted in by a funnel shift)  ==  `f > thresh`
   (R/pytocr/postprocess/db_postprocess.py:46 `segmentation = pred > self.thresh`) for every finite float32, signed zeros
   and denormal differences included (nvcc's default keeps denormals, as numpy does);
2. the row sums add the raw bit patterns of `f + 1.0f` modulo 2^32 and take the constant off once:
   sum(bits(f_i + 1) ) - n * bits(1.0)  ==  sum(round(f_i * 2^23))  for f_i in [0, 1], and the running min / max of the
   patterns flags exactly the values outside [0, 1] (NaN and Inf included) that `bits(f + 1) - bits(1) > 2^23` flags.
"""
import numpy as np


def _specials():
    th = np.float32(0.3)
    return np.array([0.0, -0.0, 1.0, th, np.nextafter(th, np.float32(1)), np.nextafter(th, np.float32(0)),
                     np.float32(1e-45), np.float32(-1e-45), np.float32(1.1754944e-38), 0.5, np.float32(0.99999994)], np.float32)


def test_sign_of_difference_is_the_threshold_test():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(200000).astype(np.float32), _specials(),
                           (rng.random(20000).astype(np.float32) - 0.5) * 4, rng.random(20000).astype(np.float32) * 1e-38])
    for th in (np.float32(0.3), np.float32(0.0), np.float32(0.5), np.float32(1e-40), np.float32(0.7), np.float32(1.0),
               np.float32(np.float16(0.3))):
        near = np.array([th, np.nextafter(th, np.float32(2)), np.nextafter(th, np.float32(-2))], np.float32)
        f = np.concatenate([vals, near])
        with np.errstate(all="ignore"):
            d = (th - f).astype(np.float32)
        assert np.array_equal(np.signbit(d), f > th)


def test_raw_bit_sums_and_range_check():
    rng = np.random.default_rng(1)
    one = np.uint32(0x3f800000)
    for n in (32, 40, 64):
        f = rng.random((5000, n)).astype(np.float32)
        f[rng.random(f.shape) < 0.02] = 1.0
        f[rng.random(f.shape) < 0.02] = 0.0
        f[rng.random(f.shape) < 0.01] = -0.0
        u = (f + np.float32(1.0)).view(np.uint32)
        want = np.rint(f.astype(np.float64) * 2 ** 23).astype(np.uint64).sum(1)       # round-half-even of f * 2^23, like f + 1.0f
        got = (u.astype(np.uint64).sum(1) & 0xffffffff) - ((n * int(one)) & 0xffffffff)
        assert np.array_equal(got & 0xffffffff, want & 0xffffffff) and want.max() < 2 ** 32
    bad = np.array([1.0000002, 1.5, 2.0, -1e-3, -1.0, np.inf, -np.inf, np.nan, 3.0e38], np.float32)
    ok = np.array([0.0, -0.0, 1.0, 0.5, -1e-12, 1e-45, np.float32(1) - np.float32(2 ** -24)], np.float32)
    for vals, flagged in ((bad, True), (ok, False)):
        with np.errstate(all="ignore"):
            u = (vals + np.float32(1.0)).view(np.uint32)
        new_rule = (u > np.uint32(0x40000000)) | (u < one)                   # running max / min of the raw patterns
        old_rule = (u - one) > np.uint32(0x800000)                           # db_scan_kernel's rule (wraps below 1.0)
        assert np.array_equal(new_rule, old_rule)
        assert new_rule.all() == flagged and new_rule.any() == flagged
