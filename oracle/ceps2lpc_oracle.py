"""CPU ORACLE (test infrastructure) for the cepstrum -> LPC step that follows Wavernn.encoder in both callers
(/root/reference/src/synthesis_qtz.py:158-160, generate_qtz_features.py:61-64).

NumPy restatement of /root/reference/src/ceps2lpc/ceps2lpc_vct.py:
  idct (:35-43), interp_band_gain (:45-57), _celt_lpc_s (:60-88), ceps2lpc_v (:122-162).
float32 arithmetic in the reference's operation order (products and sums rounded separately); the
autocorrelation (torch.fft.irfft at :137) is evaluated as the real inverse DFT in float64 and rounded to
float32, so it agrees with torch's float32 FFT to ~1e-7 relative.  Pinned against golden vectors made by the
unmodified reference (tests/golden/ceps2lpc.npz, oracle/gen_golden.py).
"""
import numpy as np

NB_BANDS, LPC_ORDER, WINDOW_SIZE, FREQ_SIZE, WINDOW_SIZE_5MS = 18, 16, 320, 161, 4
EBAND5MS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40]
COMPENSATION = np.array([0.8, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.666667, 0.5, 0.5, 0.5, 0.333333, 0.25, 0.25, 0.2,
                         0.166667, 0.173913], dtype=np.float32)


def dct_table():
    t = np.zeros((NB_BANDS, NB_BANDS), np.float32)
    for i in range(NB_BANDS):
        for j in range(NB_BANDS):
            t[i, j] = np.cos(np.float32((i + .5) * j * np.pi / NB_BANDS))
            if j == 0:
                t[i, j] = t[i, j] * np.sqrt(np.float32(.5))
    return t


def ceps2lpc(cepstrum):
    """(N, >=18) -> lpc (N,16) float32, error (N,) float32, rc (N,16) float64 (zeros after an early exit)."""
    f32 = np.float32
    c = np.asarray(cepstrum, dtype=f32)[:, :NB_BANDS].copy()
    c[:, 0] = c[:, 0] + f32(4)
    D = dct_table()
    n = len(c)
    ex = np.zeros((n, NB_BANDS), f32)
    k = np.sqrt(f32(2. / NB_BANDS)).astype(f32)
    for i in range(NB_BANDS):
        sm = np.zeros(n, f32)
        for j in range(NB_BANDS):
            sm = (sm + (c[:, j] * D[i, j]).astype(f32)).astype(f32)
        ex[:, i] = sm * k
    ex = (np.power(f32(10.0), ex).astype(f32) * COMPENSATION).astype(f32)
    g = np.zeros((n, FREQ_SIZE), f32)
    for i in range(NB_BANDS - 1):
        bs = (EBAND5MS[i + 1] - EBAND5MS[i]) * WINDOW_SIZE_5MS
        for j in range(bs):
            frac = float(j) / bs
            g[:, EBAND5MS[i] * WINDOW_SIZE_5MS + j] = (f32(1 - frac) * ex[:, i] + f32(frac) * ex[:, i + 1]).astype(f32)
    acr = np.fft.irfft(g.astype(np.float64), n=WINDOW_SIZE, axis=1)[:, :LPC_ORDER + 1].astype(f32)
    acr[:, 0] = acr[:, 0] + (acr[:, 0] * f32(0.0001) + f32(320 / 12 / 38.))
    for i in range(1, LPC_ORDER + 1):
        acr[:, i] = acr[:, i] * f32(1 - 0.00006 * i * i)
    lpc = np.zeros((n, LPC_ORDER), f32)
    err = np.zeros(n, f32)
    rc = np.zeros((n, LPC_ORDER), np.float64)
    for l in range(n):
        ac = acr[l]
        error = ac[0]
        lp = np.zeros(LPC_ORDER, f32)
        if ac[0] != 0:
            for i in range(LPC_ORDER):
                rr = f32(0.)
                for j in range(i):
                    rr = f32(rr + f32(lp[j] * ac[i - j]))
                rr = f32(rr + ac[i + 1])
                r = f32(-rr / error)
                rc[l, i] = r
                lp[i] = r
                for j in range((i + 1) // 2):
                    t1, t2 = lp[j], lp[i - 1 - j]
                    lp[j] = f32(t1 + f32(r * t2))
                    lp[i - 1 - j] = f32(t2 + f32(r * t1))
                error = f32(error - f32(f32(r * r) * error))
                if error < f32(ac[0] / f32(2 ** 10)):
                    break
                if error < f32(f32(0.001) * ac[0]):
                    break
        lpc[l] = lp
        err[l] = error
    return lpc, err, rc


def ceps2lpc_f64(cepstrum):
    """The same algorithm evaluated in float64 from the float32 input -- a GROUND TRUTH for judging float32 results
    against one another (the reference's torch float32 evaluation, this file's float32 restatement, the CUDA kernel):
    no float32 evaluation order is "the" right one, and the Levinson recursion amplifies their 1e-7 differences to
    1e-3 in the coefficients.  Table constants are the reference's float32 ones (the DCT table and the compensation
    gains are float32 tensors there, ceps2lpc_vct.py:16-33); the early exits (:82-85) are taken on the float64 error.
    -> lpc (N,16), error (N,), rc (N,16), all float64."""
    c = np.asarray(cepstrum, dtype=np.float32)[:, :NB_BANDS].astype(np.float64)
    c[:, 0] += 4.0
    D = dct_table().astype(np.float64)
    ex = (c @ D.T) * float(np.sqrt(np.float32(2. / NB_BANDS)))
    ex = np.power(10.0, ex) * COMPENSATION.astype(np.float64)
    n = len(c)
    g = np.zeros((n, FREQ_SIZE))
    for i in range(NB_BANDS - 1):
        bs = (EBAND5MS[i + 1] - EBAND5MS[i]) * WINDOW_SIZE_5MS
        for j in range(bs):
            frac = float(j) / bs
            g[:, EBAND5MS[i] * WINDOW_SIZE_5MS + j] = (1 - frac) * ex[:, i] + frac * ex[:, i + 1]
    acr = np.fft.irfft(g, n=WINDOW_SIZE, axis=1)[:, :LPC_ORDER + 1]
    acr[:, 0] += acr[:, 0] * float(np.float32(0.0001)) + float(np.float32(320 / 12 / 38.))
    for i in range(1, LPC_ORDER + 1):
        acr[:, i] *= 1 - 0.00006 * i * i
    lpc = np.zeros((n, LPC_ORDER))
    err = np.zeros(n)
    rc = np.zeros((n, LPC_ORDER))
    for l in range(n):
        ac = acr[l]
        error = ac[0]
        lp = np.zeros(LPC_ORDER)
        if ac[0] != 0:
            for i in range(LPC_ORDER):
                r = -(np.dot(lp[:i], ac[i:0:-1]) + ac[i + 1]) / error
                rc[l, i] = r
                head = lp[:i].copy()
                lp[:i] = head + r * head[::-1]
                lp[i] = r
                error = error - r * r * error
                if error < ac[0] / 2 ** 10 or error < 0.001 * ac[0]:
                    break
        lpc[l] = lp
        err[l] = error
    return lpc, err, rc
