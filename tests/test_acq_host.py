"""Host-side checks of the acquisition path that need no GPU: the classification of Doppler bins into 1-kHz classes
(`gr_acq_classify_bins`, what `gr_acq_plan_create` applies) and a numpy restatement, thread by thread, of the data flow of
the inverse acquisition kernel's packed transform (csrc/gr_acq.cu `acq_inv_kernel`, csrc/gr_fft2048t.cuh, csrc/gr_cpk.cuh):
which thread holds which element at every stage and which twiddle it applies on the input side of the next stage.  The
hardware-specific part (the tensor-memory fragment layout of the second transpose) is replaced by the transposition it
implements; everything else is the kernel's own index arithmetic."""
from __future__ import annotations

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from gps_sdr_receiver_b200 import _build
    _build.build()
    return None


def _classify(lib, bins, share=1):
    """(number of base spectra or -1, base index per bin, shift per bin, base frequencies) through the Python mirror."""
    from gps_sdr_receiver_b200 import _capi
    from gps_sdr_receiver_b200.acquisition import classify_bins
    try:
        base, shift, base_hz = classify_bins(bins, share=bool(share))
    except _capi.GrError:
        return -1, None, None, None
    return len(base_hz), base, shift, base_hz


def test_bin_classes_of_the_baseline_grids(lib):
    cold = [-10000.0 + 500.0 * b for b in range(41)]
    n, base, shift, base_hz = _classify(lib, cold)
    assert n == 2 and sorted(base_hz.tolist()) == [-500.0, 0.0]
    fine = [-10000.0 + 50.0 * b for b in range(401)]
    n, base, shift, base_hz = _classify(lib, fine)
    assert n == 20 and base_hz.min() >= -500.0 and base_hz.max() < 500.0
    # every bin is its base frequency plus shift x 1 kHz (shift taken modulo 2048)
    for bins in (cold, fine):
        n, base, shift, base_hz = _classify(lib, bins)
        q = np.where(shift >= 1024, shift - 2048, shift)
        np.testing.assert_allclose(base_hz[base] + 1000.0 * q, bins, atol=1e-9)
    # reference sweep grid (200 Hz steps): 5 classes; off-lattice spacing: one class per bin; sharing off: one per bin
    assert _classify(lib, [-5000.0 + 200.0 * b for b in range(50)])[0] == 5
    assert _classify(lib, [-3000.0 + 770.0 * b for b in range(9)])[0] == 9
    assert _classify(lib, cold, share=0)[0] == 41 and not _classify(lib, cold, share=0)[2].any()


def test_bin_class_does_not_depend_on_the_other_bins(lib):
    """What keeps a bin-sharded search bit-identical to the unsharded one: base frequency and shift of a bin are the same
    whichever subset of the grid a plan holds."""
    fine = [-10000.0 + 50.0 * b for b in range(401)]
    n, base, shift, base_hz = _classify(lib, fine)
    for lo, hi in ((0, 201), (201, 401), (37, 38), (100, 333)):
        n2, base2, shift2, base_hz2 = _classify(lib, fine[lo:hi])
        assert np.array_equal(shift2, shift[lo:hi])
        np.testing.assert_array_equal(base_hz2[base2], base_hz[base[lo:hi]])
    assert _classify(lib, [1.1e6])[0] < 0                       # beyond fs / 2


def _dft(x, axis):
    n = x.shape[axis]
    k = np.arange(n)
    w = np.exp(-2j * np.pi * np.outer(k, k) / n)
    return np.moveaxis(np.tensordot(w, np.moveaxis(x, axis, 0), axes=(1, 0)), 0, axis)


def test_inverse_transform_data_flow_matches_the_circular_correlation():
    """acq_inv_kernel's transform, restated: stage 1 "data = (C.im, C.re), twiddle = conj(X)" radix-16 over the register
    index; exchange 1; stage 2 with W_2048^((8 n2 + n3) k1) on its inputs; exchange 2; stage 3 with W_128^(n3 k2) on its
    inputs; outputs of thread t at lags fftt_out_base(t) + 128 j.  |result| must be N |ifft(X conj(C_code))|."""
    rng = np.random.default_rng(5)
    N = 2048
    X = rng.standard_normal(N) + 1j * rng.standard_normal(N)           # forward spectrum (already rotated)
    c = rng.standard_normal(N) + 1j * rng.standard_normal(N)           # stored conjugate code spectrum
    t = np.arange(128)
    j = np.arange(16)
    elem = t[:, None] + 128 * j[None, :]                               # thread t, register j <-> element t + 128 j
    # stage 1: a = (c.im, c.re) read as a complex number, times conj(X): the operand (Im Y, Re Y) of the swap-form inverse
    a = c[elem].imag + 1j * c[elem].real
    u = a * np.conj(X[elem])
    Y = X * c
    np.testing.assert_allclose(u, (Y.imag + 1j * Y.real)[elem], rtol=1e-12)
    s1 = _dft(u, axis=1)                                               # [t1][k1], t1 = 8 n2 + n3
    # exchange 1 + stage 2: thread (warp w, lane L): k1 = 4 w + 2 (L >> 4) + (L & 1), n3 = (L >> 1) & 7
    w, L = t >> 5, t & 31
    k1 = 4 * w + 2 * (L >> 4) + (L & 1)
    n3 = (L >> 1) & 7
    n2 = np.arange(16)
    src = 8 * n2[None, :] + n3[:, None]                                # writer thread of input n2
    tw2 = np.exp(-2j * np.pi * ((8 * n2[None, :] + n3[:, None]) * k1[:, None]) / 2048)
    s2 = _dft(s1[src, k1[:, None]] * tw2, axis=1)                      # [thread][k2]
    # exchange 2 + stage 3: lane L holds groups (k1loc, k2 = k2lo + 8 h) of its warp, inputs n3 = 0..7
    k1loc = 2 * (L & 1) + ((L >> 3) & 1)
    k2lo = 4 * ((L >> 2) & 1) + 2 * ((L >> 4) & 1) + ((L >> 1) & 1)
    owner = {(int(k1[i]), int(n3[i])): i for i in range(128)}           # stage-2 thread of (k1, n3)
    assert len(owner) == 128
    out = np.zeros(N, dtype=complex)
    seen = np.zeros(N, dtype=int)
    for i in range(128):
        kk1 = 4 * int(w[i]) + int(k1loc[i])
        for h in range(2):
            k2 = int(k2lo[i]) + 8 * h
            ins = np.array([s2[owner[(kk1, m)], k2] for m in range(8)])
            ins = ins * np.exp(-2j * np.pi * np.arange(8) * k2 / 128)
            res = _dft(ins[None, :], axis=1)[0]                        # k3 = 0..7
            base = 4 * int(w[i]) + int(k1loc[i]) + 16 * int(k2lo[i])   # fftt_out_base(t)
            for k3 in range(8):
                lag = base + 128 * (2 * k3 + h)
                assert lag == kk1 + 16 * k2 + 256 * k3
                out[lag] = res[k3]
                seen[lag] += 1
    assert (seen == 1).all()
    ref = N * np.fft.ifft(Y)
    np.testing.assert_allclose(np.abs(out), np.abs(ref), rtol=1e-9)
    np.testing.assert_allclose(out, ref.imag + 1j * ref.real, rtol=1e-9, atol=1e-9)   # fft(swap(y)) = swap(N ifft(y))


def test_tmem_transpose_rounds_deliver_the_radix8_groups():
    """The second transpose of the FFT goes through tensor memory (csrc/gr_fft2048t.cuh).  With the fragment layout measured
    by tools/ubench/tmem_probe.cu (profiles/tmem_probe_r01.log) --
        tcgen05.st.16x256b.x4, thread T, register 16 b + 4 i + 2 r + q  ->  lane 16 b + 8 r + T / 4, column 8 i + 2 (T % 4) + q
        tcgen05.ld.32x32b,     thread L, register c                     <-  lane L, column c
    -- two store / load rounds with the kernel's register order must leave, in lane L, the eight n3 inputs of the two
    radix-8 groups (k1loc(L), k2lo(L) + 8 h) in complex registers 8 h + 4 (n3 & 1) + 2 (n3 >> 2) + ((n3 >> 1) & 1)."""
    def one_round(regs):                                    # regs[T][rho] -> new[L][c]
        new = [[None] * 32 for _ in range(32)]
        for T in range(32):
            for rho in range(32):
                b, i, r, q = rho >> 4, (rho >> 2) & 3, (rho >> 1) & 1, rho & 1
                lane, col = 16 * b + 8 * r + T // 4, 8 * i + 2 * (T % 4) + q
                assert new[lane][col] is None
                new[lane][col] = regs[T][rho]
        return new

    regs = [[None] * 32 for _ in range(32)]
    for T in range(32):                                     # stage-2 thread: n3 = (T >> 1) & 7, k1loc = 2 (T >> 4) + (T & 1)
        n3, k1loc = (T >> 1) & 7, 2 * (T >> 4) + (T & 1)
        for k2 in range(16):
            K = 8 * ((k2 >> 2) & 1) + 4 * ((k2 >> 1) & 1) + 2 * ((k2 >> 3) & 1) + (k2 & 1)     # (e2 e1 e3 e0)
            regs[T][2 * K] = (k1loc, n3, k2, "re")
            regs[T][2 * K + 1] = (k1loc, n3, k2, "im")
    regs = one_round(one_round(regs))
    for L in range(32):
        k1loc = 2 * (L & 1) + ((L >> 3) & 1)
        k2lo = 4 * ((L >> 2) & 1) + 2 * ((L >> 4) & 1) + ((L >> 1) & 1)
        for h in range(2):
            for n3 in range(8):
                K = 8 * h + 4 * (n3 & 1) + 2 * (n3 >> 2) + ((n3 >> 1) & 1)
                assert regs[L][2 * K] == (k1loc, n3, k2lo + 8 * h, "re"), (L, h, n3, regs[L][2 * K])
                assert regs[L][2 * K + 1] == (k1loc, n3, k2lo + 8 * h, "im")


def test_shared_spectra_identity_against_the_oracle():
    """The identity behind the shared forward spectra, checked against the oracle (the reference's per-bin float32 wipe-off,
    gpsrecv.py:232-235) on the CPU: the coherent spectrum of bin f0 + q kHz is the spectrum of bin f0 rotated by q FFT
    bins, up to a constant phase, for 1-ms and for 10-ms coherent intervals; the cells of the rotated form agree with the
    oracle's grid within the magnitude tolerance."""
    from oracle import gps_oracle as orc
    from gps_sdr_receiver_b200 import synth
    sats = [synth.Sat(prn=9, doppler=3270.0, delay=700.3, amp=0.09), synth.Sat(prn=23, doppler=-1240.0, delay=50.8, amp=0.09)]
    for tcoh, k in ((1, 4), (10, 2)):
        n_ms = tcoh * k
        data = orc.raw_to_complex(synth.make_iq(sats, n_ms, seed=8))
        n = n_ms * 2048
        t = orc.sec_time(n)
        f_base = -250.0                                              # canonical member of its class
        xb, _ = orc.wipeoff(data, f_base, 0, n, t)
        base = [orc.coherent_spectrum(xb, i * tcoh, tcoh) for i in range(k)]
        for q in (-7, 0, 3, 9):
            xq, _ = orc.wipeoff(data, f_base + 1000.0 * q, 0, n, t)
            for i in range(k):
                direct = orc.coherent_spectrum(xq, i * tcoh, tcoh)
                rot = np.roll(base[i], -q)                           # rot[m] = base[m + q]
                ph = np.vdot(rot, direct) / np.vdot(rot, rot)        # the constant phase factor
                assert abs(abs(ph) - 1.0) < 1e-5
                # what is left is the reference's own float32 rounding of the phase argument w t (it grows with |f| t:
                # 6e-6 of the largest line at 9 kHz x 4 ms, 3e-5 at 9 kHz x 20 ms); the rotated form has the smaller arguments
                assert np.abs(direct - ph * rot).max() < 1e-4 * np.abs(direct).max()
        # cells: oracle grid at f_base + 1000 q vs statistics of the rotated base spectra
        ref = orc.acq_grid(data, [9, 23], f_base + 3000.0, 1000.0, 1, tcoh, k, orc.ACQ_MODE_POW)
        for i, p in enumerate((9, 23)):
            cs = np.conjugate(orc.code_spectrum(p))
            stat = np.zeros(2048)
            for s in base:
                c = np.fft.ifft(np.roll(s, -3) * cs)
                stat += c.real ** 2 + c.imag ** 2
            st = orc.acq_cell_stats(stat)
            assert int(st["mx"]) == int(ref["mx"][i, 0])
            for key in ("peak", "mean", "std", "z"):
                assert abs(st[key] - ref[key][i, 0]) <= 1e-4 * abs(ref[key][i, 0]), (tcoh, p, key)
