"""Independent numpy float64 restatement of SURVEY.md Appendix A in its v0 form (explicit
atan2 phases + rint wrap), used only to pin the C oracle (oracle/pv_ref.c).  Small inputs only."""
import numpy as np


def pv_numpy(x, N, H, r, fs=48000.0):
    x = np.asarray(x, np.float32)
    n = x.size
    F = (n + H - 1) // H
    M = N // 2
    nb = M + 1
    osamp = N // H
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    g = np.float32(H / np.sum(w.astype(np.float64) ** 2))
    xp = np.concatenate([np.zeros(N, np.float32), x, np.zeros(N, np.float32)])
    kk = np.arange(nb)
    jk = np.trunc(kk.astype(np.float32) * np.float32(r)).astype(np.int64)
    phi_prev = np.zeros(nb)
    acc = np.zeros(nb, np.uint64)
    out = np.zeros(n + 2 * N)
    kmin = max(1, int(np.ceil(50 * N / fs)))
    kmax = min(M, int(np.floor(2000 * N / fs)))
    peak = np.zeros(F, np.int32)
    f0 = np.zeros(F)
    for f in range(F):
        s0 = (f + 1) * H - N
        fr = xp[s0 + N:s0 + 2 * N].astype(np.float64) * w.astype(np.float64)
        X = np.fft.rfft(fr)
        mag = np.abs(X)
        phi = np.angle(X)
        d = phi - phi_prev - 2 * np.pi * kk / osamp
        d = d - 2 * np.pi * np.rint(d / (2 * np.pi))
        d[0] = abs(d[0])   # PV-spec v1: the purely real bins sit on the cut; Im := +0 -> d in {0, +pi}
        d[M] = abs(d[M])
        phi_prev = phi
        nu = kk + osamp * d / (2 * np.pi)
        pk = kmin + int(np.argmax(mag[kmin:kmax + 1]))
        peak[f] = pk
        f0[f] = nu[pk] * fs / N
        smag = np.zeros(nb)
        snu = kk.astype(np.float64).copy()
        for k in range(nb):
            j = jk[k]
            if 0 <= j < nb:
                smag[j] += mag[k]
                snu[j] = np.float64(np.float32(r)) * nu[k]
        t = snu / osamp
        t = t - np.floor(t)
        inc = np.rint(t * 4294967296.0).astype(np.uint64) % (1 << 32)
        acc = (acc + inc) % (1 << 32)
        th = acc.astype(np.float64) * (2 * np.pi / 4294967296.0)
        Y = smag * np.exp(1j * th)
        Y[0] = Y[0].real
        Y[M] = Y[M].real
        y = np.fft.irfft(Y, N)
        out[s0 + N:s0 + 2 * N] += np.float64(g) * w * y
    return out[N:N + n].astype(np.float32), peak, f0
