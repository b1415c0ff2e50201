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


K1_OF = [0, 1, 4, 7, 2, 3, 6, 5]          # mel.cu k1_of(): k1 handled by threads with (thread >> 3) == h
H_OF = [K1_OF.index(k) for k in range(8)]  # mel.cu h_of()


def kernel_fft1024_real(xw):
    """xw: 1024 windowed real samples -> X[0..512], following the kernel's thread / shared-memory / shuffle
    choreography: pass 1 -> shared memory -> pass 2 -> shared memory (same buffer) -> pass 3 -> partner shuffles."""
    tw = np.exp(-2j * np.pi * np.arange(1024) / 1024)                     # device twiddle table
    z = xw[0::2] + 1j * xw[1::2]                                           # 512 complex points
    S1 = np.zeros(8 * 72, complex)
    for n2 in range(64):                                                   # pass 1: thread n2
        y = dft8([z[64 * n1 + n2] for n1 in range(8)])
        for k1 in range(8):
            S1[k1 * 72 + n2] = y[k1] * tw[2 * ((n2 * k1) % 512)]           # W_512^(n2*k1)
    regs = [[0j] * 8 for _ in range(64)]
    S2 = np.zeros(8 * 72, complex)                                         # the same buffer, after a group barrier
    for t in range(64):                                                    # pass 2: thread (h, b), k1 = K1_OF[h]
        h, b = t // 8, t % 8
        k1 = K1_OF[h]
        u = dft8([S1[k1 * 72 + 8 * a + b] for a in range(8)])
        for c in range(8):
            S2[h * 72 + c * 9 + b] = u[c] * tw[16 * ((b * c) % 64)]        # W_64^(b*c): U[k1][c][b]
    for t in range(64):                                                    # pass 3: thread (h, c): regs[t][d] = Z[k1+8c+64d]
        h, c = t // 8, t % 8
        regs[t] = dft8([S2[h * 72 + c * 9 + b] for b in range(8)])
    X = np.zeros(513, complex)
    w16 = np.exp(-2j * np.pi * np.arange(8) / 16)
    for t in range(64):                                                    # recombination with the partner thread
        h, c = t // 8, t % 8
        k1 = K1_OF[h]
        if k1 == 0 and c == 0:
            for d in range(8):
                za, zb = regs[t][d], np.conj(regs[t][(8 - d) % 8])
                ze, zo = (za + zb) / 2, (za - zb) / 2j
                X[64 * d] = ze + w16[d] * zo
                if d == 0:
                    X[512] = ze - w16[d] * zo
            continue
        k1p, cp = (8 - k1) % 8, (7 - c) if k1 else (8 - c) % 8
        hp = H_OF[k1p]
        assert hp // 4 == h // 4, "partner threads must share a warp"
        lane_p = 8 * (hp % 4) + cp
        tp = 32 * (t // 32) + lane_p
        wbase = tw[k1 + 8 * c]
        for dd in range(8):
            d = 7 - dd
            za, zb = regs[t][d], np.conj(regs[tp][dd])
            ze, zo = (za + zb) / 2, (za - zb) / 2j
            X[k1 + 8 * c + 64 * d] = ze + wbase * w16[d] * zo
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


def test_smem_maps_are_injective_and_conflict_free():
    s1 = {k1 * 72 + n2 for k1 in range(8) for n2 in range(64)}
    s2 = {h * 72 + c * 9 + b for h in range(8) for c in range(8) for b in range(8)}
    assert len(s1) == 512 and max(s1) < 576 and len(s2) == 512 and max(s2) < 576
    # pass 2 reads S1[k1 * 72 + 8 q + lo] as 8-byte words: a half-warp (16 lanes = two h values) must hit 16 distinct
    # bank pairs for every q (the k1 permutation keeps the two rows an odd number of rows apart)
    for half in range(4):
        for q in range(8):
            banks = {((K1_OF[2 * half + hh] * 72 + 8 * q + lo) * 2) % 32 for hh in range(2) for lo in range(8)}
            assert len(banks) == 16
    # magnitude stores mag[k], k = k1 + 8 c + 64 d, for the 32 lanes of a warp and a fixed d: at most 2-way (c and c + 4
    # share a bank: 16 distinct banks, 8 such stores per frame and warp)
    for w in range(2):
        for d in range(8):
            banks = [(K1_OF[4 * w + hh] + 8 * c + 64 * d) % 32 for hh in range(4) for c in range(8)]
            assert len(set(banks)) == 16 and max(banks.count(x) for x in set(banks)) == 2


def test_partner_threads_share_a_warp_and_cover_every_bin():
    seen = set()
    for t in range(64):
        h, c = t // 8, t % 8
        k1 = K1_OF[h]
        k1p = (8 - k1) % 8
        assert H_OF[k1p] // 4 == h // 4
        for d in range(8):
            seen.add(k1 + 8 * c + 64 * d)
    assert seen == set(range(512))
    assert sorted(K1_OF) == list(range(8)) and all(K1_OF[H_OF[k]] == k for k in range(8))
    # the packed tables in mel.cu
    assert [(0x56327410 >> (4 * h)) & 7 for h in range(8)] == K1_OF
    assert [(0x36725410 >> (4 * k)) & 7 for k in range(8)] == H_OF


def test_parseval_energy_identity():
    """mel.cu computes sum_{k<=512} |X_k|^2 as (1024 * sum xw^2 + X_0^2 + X_512^2) / 2 (stft.py:84 sums the bins)."""
    rng = np.random.default_rng(2)
    for _ in range(3):
        x = rng.uniform(-1, 1, 1024) * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024))
        X = np.fft.rfft(x)
        lhs = (np.abs(X) ** 2).sum()
        rhs = 0.5 * (1024.0 * (x ** 2).sum() + X[0].real ** 2 + X[512].real ** 2)
        np.testing.assert_allclose(lhs, rhs, rtol=1e-12)


def test_filterbank_span_form_is_exact():
    """Mirror of e2e_mel_create's sparse filterbank (mel.cu): per filter the span [lo, hi] of its non-zero bins with
    the weights packed back to back; the kernel's per-(filter, frame) dot product over the span reproduces basis @ mag.
    The magnitude tile's rows are whole float4s with a pitch of 4 x odd words (conflict-free 128-bit lane reads)."""
    from oracle import mel_oracle as mo
    basis = mo.slaney_mel_basis().astype(np.float64)
    n_mels, nbins = basis.shape
    lo = [int(np.nonzero(basis[r])[0][0]) for r in range(n_mels)]
    hi = [int(np.nonzero(basis[r])[0][-1]) for r in range(n_mels)]
    nb = 1 + max(hi)
    assert nb == 372
    off, wpk = [], []
    for r in range(n_mels):
        off.append(len(wpk))
        wpk.extend(basis[r, lo[r]: hi[r] + 1])
    assert len(wpk) <= 1024                                # a few KB of shared memory per CTA
    rng = np.random.default_rng(3)
    mag = rng.uniform(0, 10, nbins)
    got = np.array([sum(wpk[off[r] + i] * mag[lo[r] + i] for i in range(hi[r] - lo[r] + 1)) for r in range(n_mels)])
    np.testing.assert_allclose(got, basis @ mag, rtol=1e-12)
    # rows are whole float4s, pitch = 4 x odd: the 8 lanes (frames) of a quarter warp read 8 distinct 16-byte bank groups
    magp = (((nb + 3) // 4) | 1) * 4
    assert magp >= nb and (magp // 4) % 2 == 1
    for k4 in (0, 5, 92):
        for q in range(4):
            assert len({(f * (magp // 4) + k4) % 8 for f in range(8 * q, 8 * q + 8)}) == 8
    # spans widened to whole groups of four bins stay inside a row
    for r in range(n_mels):
        lo4 = lo[r] & ~3
        assert lo4 + (hi[r] + 1 - lo4 + 3) // 4 * 4 <= magp
