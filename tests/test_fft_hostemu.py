"""Index logic of the CUDA Fourier and table kernels, run on the CPU through the test-only harness
tests/hostemu (same __host__ __device__ code the kernels call)."""
import ctypes

import numpy as np
import pytest

import ectrans_oracle as eo

DP = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(DP)


def _smooth_lengths():
    out = []
    for n in list(range(2, 400, 2)) + [1024, 2310, 3596, 4096, 4 * 13 * 17, 7776, 10368]:
        out.append(n)
    return out


def test_fft_all_small_even_lengths(emu):
    rng = np.random.default_rng(0)
    checked = 0
    for n in _smooth_lengths():
        if not emu.emu_smooth(n):
            continue
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        xin = np.ascontiguousarray(np.stack([x.real, x.imag], 1))
        out = np.zeros((n, 2))
        assert emu.emu_fft(n, P(xin), P(out)) == 0
        ref = np.fft.ifft(x) * n
        assert np.abs(out[:, 0] + 1j * out[:, 1] - ref).max() <= 5e-15 * np.abs(ref).max() * max(1, np.log2(n))
        checked += 1
    assert checked > 120


def test_fft_float_core(emu):
    """The same templated core instantiated on float2 (sp handles): smooth lengths directly, and the
    chirp-z chain (DIF stages, fused middle, DIT stages) with a unit kernel, which must return n * swap(x)."""
    rng = np.random.default_rng(1)
    for n in (20, 48, 336, 1024, 1616, 2310, 4096, 10368, 16384):
        if not emu.emu_smooth(n):
            continue
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        xin = np.ascontiguousarray(np.stack([x.real, x.imag], 1))
        out = np.zeros((n, 2))
        assert emu.emu_fft_float(n, P(xin), P(out), 0) == 0
        ref = np.fft.ifft(x) * n
        err = np.linalg.norm(out[:, 0] + 1j * out[:, 1] - ref) / np.linalg.norm(ref)
        assert err < 1e-6, (n, err)
    for n in (64, 1024, 3 * 512, 5 * 1024, 7 * 2048, 16384):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        xin = np.ascontiguousarray(np.stack([x.real, x.imag], 1))
        out = np.zeros((n, 2))
        assert emu.emu_fft_float(n, P(xin), P(out), 1) == 0
        # F+(swap(F+(x))) = i conj(F-(F+(x))) = n swap(x)   (swap(z) = i conj(z); the kernels swap on load)
        want = n * (x.imag + 1j * x.real)
        err = np.linalg.norm(out[:, 0] + 1j * out[:, 1] - want) / np.linalg.norm(want)
        assert err < 2e-6, (n, err)


def _pair(emu, nlon, km, force, nthr, rng):
    sa = rng.standard_normal(km + 1) + 1j * rng.standard_normal(km + 1)
    sb = rng.standard_normal(km + 1) + 1j * rng.standard_normal(km + 1)
    spec = np.ascontiguousarray(np.stack([sa.real, sa.imag, sb.real, sb.imag], 1))
    oa, ob = np.zeros(nlon), np.zeros(nlon)
    blue = emu.emu_ftinv_pair(nlon, km, P(spec), P(oa), P(ob), nthr, force)

    def ref(s):
        h = np.zeros(nlon // 2 + 1, complex)
        h[:km + 1] = s
        h[0] = h[0].real
        return np.fft.irfft(h, n=nlon) * nlon

    e1 = max(np.abs(oa - ref(sa)).max(), np.abs(ob - ref(sb)).max()) / np.abs(ref(sa)).max()
    ra, rb = rng.standard_normal(nlon), rng.standard_normal(nlon)
    sp = np.zeros((km + 1, 4))
    emu.emu_ftdir_pair(nlon, km, P(ra), P(rb), P(sp), nthr, force)
    fa, fb = np.fft.rfft(ra)[:km + 1] / nlon, np.fft.rfft(rb)[:km + 1] / nlon
    e2 = max(np.abs(sp[:, 0] + 1j * sp[:, 1] - fa).max(), np.abs(sp[:, 2] + 1j * sp[:, 3] - fb).max()) / np.abs(fa).max()
    return blue, e1, e2


@pytest.mark.parametrize("nlon,km", [(20, 8), (24, 10), (18, 8), (30, 14), (300, 148), (336, 79), (656, 159),
                                     (5136, 1279), (4 * 1283, 1279), (148, 40), (148, 73), (134, 50), (212, 60),
                                     (20 + 4 * 399, 399), (4 * 97, 120)])
def test_pair_transforms(emu, nlon, km):
    rng = np.random.default_rng(nlon)
    blue, e1, e2 = _pair(emu, nlon, km, 0, 7, rng)
    assert e1 < 5e-15 and e2 < 5e-15, (blue, e1, e2)


@pytest.mark.parametrize("nlon,km", [(25, 12), (27, 13), (45, 20), (75, 37), (81, 40), (125, 62), (135, 40), (243, 121),
                                     (375, 100), (1125, 562), (3, 1), (5, 2)])
def test_pair_transforms_odd_lengths(emu, nlon, km):
    """Odd row lengths (classic reduced Gaussian grids: 25, 27, 45, 75, 81, 125, ...) always go through chirp-z; the
    reflected chirp entries change sign there (c[N - j] = (-1)^N c[j])."""
    rng = np.random.default_rng(nlon)
    blue, e1, e2 = _pair(emu, nlon, km, 0, 7, rng)
    assert blue == 1 and e1 < 5e-15 and e2 < 5e-15, (blue, e1, e2)


@pytest.mark.parametrize("nlon,km", [(20, 8), (24, 11), (18, 8), (300, 148), (336, 79)])
def test_pair_transforms_forced_chirpz(emu, nlon, km):
    rng = np.random.default_rng(nlon + 1)
    blue, e1, e2 = _pair(emu, nlon, km, 1, 32, rng)
    assert blue == 1 and e1 < 5e-15 and e2 < 5e-15


def test_every_octahedral_length_has_a_plan(emu):
    # O1280: nlon = 20 + 4 i ; lengths that are not <=31-smooth go through chirp-z, which must fit shared memory
    rng = np.random.default_rng(5)
    for i in (0, 1, 7, 330, 331, 777, 1000, 1279):
        nlon = 20 + 4 * i
        km = min(1279, max(2, nlon // 3 - 1))
        blue, e1, e2 = _pair(emu, nlon, km, 0, 5, rng)
        assert e1 < 1e-14 and e2 < 1e-14


@pytest.mark.parametrize("T,N,ms", [(79, 80, None), (399, 400, [0, 1, 2, 3, 57, 200, 398, 399]),
                                    (1279, 1280, [0, 1, 2, 641, 1277, 1278, 1279])])
def test_supolf_column_bitwise(emu, T, N, ms):
    """The table kernel's recurrence reproduces the oracle's SUPOLF restatement bit for bit on the CPU."""
    mu, _ = eo.gauss_latitudes(2 * N)
    mun = np.ascontiguousarray(mu[:N])
    for m in (range(T + 1) if ms is None else ms):
        imaxn = T + 1
        ila, ils = (T - m + 2) // 2, (T - m + 3) // 2
        even = (imaxn - m) % 2 == 0
        for par, kc, kn in ((1, ila, imaxn + 1 if even else imaxn), (0, ils, imaxn if even else imaxn + 1)):
            if kc == 0:
                continue
            ref = eo.supolf(m, kn, mun, kcheap=3 if par else 2)[m + par:m + par + 2 * kc:2]
            out = np.zeros((kc, N))
            emu.emu_supolf(m, par, kc, kn, N, P(mun), P(out))
            assert np.array_equal(out, ref), (m, par)


def test_every_row_of_o1280_and_every_small_length(emu):
    """Every row length of the TCo1279 / O1280 benchmark grid with its own NMEN, and every length 3 .. 600 (odd and even)
    at the largest truncation it can carry: the plan builder finds a plan (smooth or chirp-z) and the pair transforms
    agree with pocketfft to a few ulp."""
    rng = np.random.default_rng(3)
    s = eo.setup(1279, 2560, eo.octahedral_nloen(1280), tables=False)
    worst, blue = 0.0, 0
    for i in range(1280):
        b, e1, e2 = _pair(emu, 20 + 4 * i, max(int(s.nmen[i]), 1), 0, 5, rng)
        worst = max(worst, e1, e2); blue += b
    assert worst < 1e-14 and blue == 774                # 60 % of the rows (68 % of the points) go through chirp-z
    for n in range(3, 601):
        b, e1, e2 = _pair(emu, n, max((n - 1) // 2, 1), 0, 3, rng)
        assert e1 < 1e-14 and e2 < 1e-14, (n, b, e1, e2)


# ---- chirp-z rows split over a CTA pair (csrc/fourier_cz.h) ----
def test_half_plans_use_one_composite_radix(emu):
    """H = r 2^k half plans: 16s innermost, the odd factor merged with the left-over power of two into one
    register-resident radix, so the four largest TCo1279 classes are three-stage plans."""
    import ctypes as C
    rad = (C.c_int * 16)()
    for n, want in ((4096, [16, 16, 16]), (3584, [16, 16, 14]), (3072, [16, 16, 12]), (2560, [16, 16, 10]),
                    (8192, [16, 16, 16, 2]), (7168, [16, 16, 2, 14]), (6144, [16, 16, 2, 12]), (5120, [16, 16, 2, 10]),
                    (1536, [16, 16, 6]), (768, [16, 16, 3]), (24, [8, 3]), (160, [16, 10]), (448, [16, 2, 14])):
        k = emu.emu_half_plan(n, rad)
        assert list(rad[:k]) == want, (n, list(rad[:k]))
        assert int(np.prod(rad[:k])) == n


def test_fft_composite_radices(emu):
    rng = np.random.default_rng(2)
    for n in (96, 160, 192, 224, 448, 1536, 2560, 3072, 3584, 5120, 6144, 7168):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        xin = np.ascontiguousarray(np.stack([x.real, x.imag], 1))
        out = np.zeros((n, 2))
        assert emu.emu_fft_half(n, P(xin), P(out)) == 0
        ref = np.fft.ifft(x) * n
        assert np.abs(out[:, 0] + 1j * out[:, 1] - ref).max() <= 5e-15 * np.abs(ref).max() * np.log2(n)


def _cz_pair(emu, nlon, km, nthr, rng):
    sa = rng.standard_normal(km + 1) + 1j * rng.standard_normal(km + 1)
    sb = rng.standard_normal(km + 1) + 1j * rng.standard_normal(km + 1)
    spec = np.ascontiguousarray(np.stack([sa.real, sa.imag, sb.real, sb.imag], 1))
    oa, ob = np.zeros(nlon), np.zeros(nlon)
    H = emu.emu_cz_inv_pair(nlon, km, P(spec), P(oa), P(ob), nthr)
    assert H > 0, H

    def ref(s):
        h = np.zeros(nlon // 2 + 1, complex)
        h[:km + 1] = s
        h[0] = h[0].real
        return np.fft.irfft(h, n=nlon) * nlon if nlon % 2 == 0 else np.fft.irfft(h, n=nlon) * nlon

    e1 = max(np.abs(oa - ref(sa)).max(), np.abs(ob - ref(sb)).max()) / np.abs(ref(sa)).max()
    ra, rb = rng.standard_normal(nlon), rng.standard_normal(nlon)
    sp = np.zeros((km + 1, 4))
    assert emu.emu_cz_dir_pair(nlon, km, P(ra), P(rb), P(sp), nthr) == H
    fa, fb = np.fft.rfft(ra)[:km + 1] / nlon, np.fft.rfft(rb)[:km + 1] / nlon
    e2 = max(np.abs(sp[:, 0] + 1j * sp[:, 1] - fa).max(), np.abs(sp[:, 2] + 1j * sp[:, 3] - fb).max()) / np.abs(fa).max()
    return H, e1, e2


@pytest.mark.parametrize("nlon,km", [(20, 8), (24, 11), (148, 40), (148, 73), (212, 60), (4 * 97, 120), (4 * 1283, 1279),
                                     (5136, 1279), (20 + 4 * 1000, 1279), (20 + 4 * 700, 1100), (100, 49), (27, 13),
                                     (75, 37), (1125, 562), (10256, 2559), (3, 1), (5, 2)])
def test_cz_pair_split(emu, nlon, km):
    """The split chirp-z (even / odd bins as two half-length convolutions, combined in the output phase) against
    pocketfft; covers N > H and N < H, odd row lengths, the TCo1279 equator and the TCo2559 equator (H = 8192)."""
    rng = np.random.default_rng(nlon + km)
    H, e1, e2 = _cz_pair(emu, nlon, km, 7, rng)
    assert e1 < 1e-14 and e2 < 1e-14, (H, e1, e2)


def test_cz_every_chirpz_row_of_o1280(emu):
    rng = np.random.default_rng(4)
    s = eo.setup(1279, 2560, eo.octahedral_nloen(1280), tables=False)
    worst, n = 0.0, 0
    for i in range(0, 1280):
        nlon = 20 + 4 * i
        if emu.emu_smooth(nlon):
            continue
        if i % 3 and i < 1270:
            continue
        H, e1, e2 = _cz_pair(emu, nlon, max(int(s.nmen[i]), 1), 5, rng)
        worst = max(worst, e1, e2); n += 1
    assert worst < 1e-14 and n > 250
