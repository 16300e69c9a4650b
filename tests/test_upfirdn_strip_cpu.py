"""Host restatement of two pieces of the upfirdn2d strip kernel (csrc/upfirdn2d.cu, upfirdn2d_blur_strip_kernel), no GPU needed:
  * the on-device rank-one factorisation of the 4x4 taps (pivot row x pivot column / pivot, accepted when it reproduces every tap to
    1e-6 of the pivot): the blur kernels of the models (stylegan2/model.py:19-27 make_kernel) factor, a general kernel does not;
  * the staging of a 16-bit row as the ALIGNED 32-bit words that cover it: a row whose first element sits at an odd index lands one
    slot to the right and the window reader shifts it back with a funnel shift; words that straddle the image border are zeroed."""
import numpy as np


def _factor(k4):
    sk = np.zeros((4, 4), np.float32)
    kh, kw = k4.shape
    sk[:kh, :kw] = k4[::-1, ::-1]                       # flipped taps (upfirdn2d_kernel.cu:77)
    piv = int(np.argmax(np.abs(sk)))
    pv = sk.flat[piv]
    inv = np.float32(1.0) / pv if pv != 0 else np.float32(0)
    kc = sk[piv >> 2, :].copy()
    kr = (sk[:, piv & 3] * inv).astype(np.float32)
    dev = np.abs(sk - np.outer(kr, kc).astype(np.float32)).max()
    return bool(dev <= np.float32(1e-6) * abs(pv)), kr, kc, sk


def test_tap_factorisation():
    k1 = np.array([1, 3, 3, 1], np.float32)
    blur = (np.outer(k1, k1) / k1.sum() ** 2).astype(np.float32)
    for k in (blur, blur * 4, blur[:3, :2].copy(), np.ones((1, 1), np.float32)):
        sep, kr, kc, sk = _factor(k)
        assert sep
        assert np.allclose(np.outer(kr, kc), sk, rtol=0, atol=1e-6 * np.abs(sk).max())
    rng = np.random.default_rng(0)
    assert not _factor(rng.standard_normal((4, 4)).astype(np.float32))[0]          # full rank: the 16-tap path
    assert _factor(np.zeros((4, 4), np.float32))[0]                                 # all-zero taps: trivially rank one (output 0)


def _stage_row_words(plane_u16, e, slots=132):
    """The 66 aligned 32-bit words that cover elements e .. e + 131 of the flat 16-bit tensor (zero where out of range)."""
    m = e & 1
    first = (e - m) // 2
    words = np.zeros(slots // 2, np.uint32)
    for w in range(slots // 2):
        idx = 2 * (first + w)
        lo = int(plane_u16[idx]) if 0 <= idx < plane_u16.size else 0
        hi = int(plane_u16[idx + 1]) if 0 <= idx + 1 < plane_u16.size else 0
        words[w] = lo | (hi << 16)
    return words, m


def _window(words, lane, m):
    """strip_window<T>: columns 4*lane .. 4*lane + 6 of the staged row from two 8-byte loads and a funnel shift by 16*m bits."""
    a0, a1, b0, b1 = (int(words[2 * lane + j]) for j in range(4))
    sh = 16 * m
    f = lambda lo, hi: ((lo | (hi << 32)) >> sh) & 0xFFFFFFFF
    w = [f(a0, a1), f(a1, b0), f(b0, b1), b1 >> sh]
    out = []
    for x in w:
        out += [x & 0xFFFF, x >> 16]
    return out[:7]


def test_sixteen_bit_row_staging_and_window():
    rng = np.random.default_rng(1)
    in_w, rows = 1025, 6                                 # odd width: rows alternate between even and odd first-element index
    t = rng.integers(1, 65535, size=rows * in_w, dtype=np.uint16)
    for r in range(rows):
        for ix0 in (0, 127, 894):                        # tile column 0 at image column ix0 (interior tiles)
            e = r * in_w + ix0
            words, m = _stage_row_words(t, e)
            assert m == (e & 1)
            for lane in (0, 1, 17, 31):
                got = _window(words, lane, m)
                want = [int(v) for v in t[e + 4 * lane: e + 4 * lane + 7]]
                assert got == want, (r, ix0, lane)


def test_border_fixup_slots():
    """A word that straddles the left / right image border brings one element of the neighbouring row: the kernel zeroes slot
    (-1 - ix0 + m) and slot (in_w - ix0 + m) of every staged row of a border tile — exactly the image columns -1 and in_w."""
    in_w = 1025
    for ix0 in (-1, -2):
        for e_par in (0, 1):
            m = e_par
            s_lo = -1 - ix0 + m
            assert s_lo - m + ix0 == -1                   # slot s holds tile column s - m = image column s - m + ix0
    ix0 = 1024 - 128 + 1                                  # last tile column of a pad-(1,1) blur: columns 897 .. 1027
    for m in (0, 1):
        s_hi = in_w - ix0 + m
        assert 0 <= s_hi < 132 and s_hi - m + ix0 == in_w
