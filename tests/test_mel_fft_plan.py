"""CPU emulation of the index math of the mel kernel's 1024-point real FFT (e2e_tts_b200/csrc/mel.cu):
real->complex packing, three radix-8 passes over 64 threads with the same shared-memory index maps and
twiddles, and the real-FFT post-pass.  Catches decomposition bugs before any GPU time is spent."""
import numpy as np


def dft8(a):
    """Radix-2 DIF 8-point DFT, natural-order output — the butterfly network of dft8() in mel.cu."""
    r = np.sqrt(0.5)
    b = [a[i] + a[i + 4] for i in range(4)] + [a[i] - a[i + 4] for i in range(4)]
    b[5] = b[5] * complex(r, -r)
    b[6] = b[6] * complex(0, -1)
    b[7] = b[7] * complex(-r, -r)

    def dft4(c):
        d0, d2 = c[0] + c[2], c[0] - c[2]
        d1, d3 = c[1] + c[3], (c[1] - c[3]) * complex(0, -1)
        return [d0 + d1, d2 + d3, d0 - d1, d2 - d3]

    e, o = dft4(b[:4]), dft4(b[4:])
    out = [0] * 8
    for m in range(4):
        out[2 * m], out[2 * m + 1] = e[m], o[m]
    return out


def kernel_fft1024_real(xw):
    """xw: 1024 windowed real samples -> X[0..512], following the kernel's thread/smem choreography."""
    tw = np.exp(-2j * np.pi * np.arange(1024) / 1024)                     # device twiddle table
    z = xw[0::2] + 1j * xw[1::2]                                           # 512 complex points
    S1 = np.zeros(8 * 72, complex)
    S2 = np.zeros(8 * 72, complex)
    S3 = np.zeros(512 + 64, complex)
    for n2 in range(64):                                                   # pass 1: thread n2
        y = dft8([z[64 * n1 + n2] for n1 in range(8)])
        for k1 in range(8):
            S1[k1 * 72 + n2] = y[k1] * tw[2 * ((n2 * k1) % 512)]           # W_512^(n2*k1)
    for t in range(64):                                                    # pass 2: thread (k1, b)
        k1, b = t // 8, t % 8
        u = dft8([S1[k1 * 72 + 8 * a + b] for a in range(8)])
        for c in range(8):
            S2[k1 * 72 + c * 9 + b] = u[c] * tw[16 * ((b * c) % 64)]       # W_64^(b*c)
    for t in range(64):                                                    # pass 3: thread (k1, c)
        k1, c = t // 8, t % 8
        v = dft8([S2[k1 * 72 + c * 9 + b] for b in range(8)])
        for d in range(8):
            k = k1 + 8 * c + 64 * d
            S3[k ^ ((k >> 3) & 7)] = v[d]
    X = np.zeros(513, complex)
    for t in range(64):                                # post pass: thread t, bin pairs (k, 512 - k), k = t + 64 j
        for j in range(4):
            k = t + 64 * j
            kk = (512 - k) % 512
            a, bq = S3[k ^ ((k >> 3) & 7)], np.conj(S3[kk ^ ((kk >> 3) & 7)])
            ze, zo = (a + bq) / 2, (a - bq) / 2j
            wz = tw[k] * zo
            X[k] = ze + wz
            X[512 - k] = np.conj(ze - wz)
    z = S3[256]                                        # thread 0: bin 256 (only |X| is used by the kernel)
    X[256] = np.conj(z)
    return X


def test_dft8_butterfly():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    np.testing.assert_allclose(dft8(list(a)), np.fft.fft(a), atol=1e-12)


def test_fft_choreography_matches_rfft():
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.uniform(-1, 1, 1024) * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024))
        np.testing.assert_allclose(kernel_fft1024_real(x), np.fft.rfft(x), atol=1e-10)


def test_smem_maps_are_injective():
    s1 = {k1 * 72 + n2 for k1 in range(8) for n2 in range(64)}
    s2 = {k1 * 72 + c * 9 + b for k1 in range(8) for c in range(8) for b in range(8)}
    s3 = {k ^ ((k >> 3) & 7) for k in range(512)}
    assert len(s1) == 512 and len(s2) == 512 and len(s3) == 512
    assert max(s1) < 576 and max(s2) < 576 and max(s3) < 576


def test_parseval_energy_identity():
    """mel.cu computes sum_{k<=512} |X_k|^2 as (1024 * sum xw^2 + X_0^2 + X_512^2) / 2 (stft.py:84 sums the bins)."""
    rng = np.random.default_rng(2)
    for _ in range(3):
        x = rng.uniform(-1, 1, 1024) * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024))
        X = np.fft.rfft(x)
        lhs = (np.abs(X) ** 2).sum()
        rhs = 0.5 * (1024.0 * (x ** 2).sum() + X[0].real ** 2 + X[512].real ** 2)
        np.testing.assert_allclose(lhs, rhs, rtol=1e-12)


def test_filterbank_block_form_is_exact_and_fits():
    """Mirror of e2e_mel_create's sparse filterbank (mel.cu): thread t of a frame owns bins [6t, 6t + 6) and the filters
    overlapping them as a dense 5 x 6 block; one partial per (filter, thread); fixed-order sums reproduce basis @ mag.
    Checks the bounds the kernel variant <BPT = 6, FPB = 5> relies on for the e2e-tts basis."""
    from oracle import mel_oracle as mo
    basis = mo.slaney_mel_basis().astype(np.float64)
    n_mels, nbins = basis.shape
    nb = 1 + max(k for k in range(nbins) if basis[:, k].any())
    assert nb == 372 and 64 * 6 >= nb                      # bins the kernel computes magnitudes for
    BPT, FPB = 6, 5
    rng = np.random.default_rng(3)
    mag = rng.uniform(0, 10, nbins)
    first, count, part = {}, {}, {}
    for t in range(64):
        bins = [k for k in range(t * BPT, (t + 1) * BPT) if k < nbins]
        filt = [r for r in range(n_mels) if basis[r, bins].any()]
        assert len(filt) <= FPB
        for r in filt:
            first.setdefault(r, t)
            count[r] = t - first[r] + 1
            part[(r, t - first[r])] = sum(basis[r, k] * mag[k] for k in bins)
    split = max(count.values())
    assert split <= 12
    got = np.array([sum(part.get((r, q), 0.0) for q in range(split)) for r in range(n_mels)])
    np.testing.assert_allclose(got, basis @ mag, rtol=1e-12)
    # padded magnitude index: pitch 7 per 6-bin block keeps the 64 threads on distinct banks
    pad = lambda k: k + (k // BPT) * ((BPT | 1) - BPT)
    assert len({pad(k) for k in range(384)}) == 384 and max(pad(k) for k in range(384)) < 584
    for i in range(BPT):
        assert len({(t * 7 + i) % 32 for t in range(32)}) == 32
