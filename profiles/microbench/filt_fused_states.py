"""Feasibility study (CPU, no GPU): can the backward filtfilt pass take its chunk states from the FORWARD
final pass, as order-independent dot products, instead of re-reading y1 in a backward local pass?

    python profiles/microbench/filt_fused_states.py

Backward pass of scipy.signal.filtfilt = lfilter over reversed y1 from zi * y1[-1].  Three evaluations:
  ref  scipy's sequential recurrence (what the reference computes)
  A    today's scheme (csrc/filtfilt.cu): 256-sample chunks counted from the START of the reversed signal,
       zero-state recurrence per chunk -> carry z' = A^256 z + s in extended precision -> exact re-run
  B    proposed: chunks aligned with the FORWARD chunk grid (so that the forward final pass, which has the
       y1 values of a chunk in registers, can emit the state), zero-state final state of a chunk formed as
       s = sum_j (A^j g) y1[p + j] - five dot products in any order - partial first chunk run sequentially
The carry is done in 50-digit arithmetic (the kernels use double-double); everything else in float64.
Prints max |y - ref| per scheme and the size of the dot-product coefficients.
"""
import numpy as np
from mpmath import mp, matrix, mpf
from scipy import signal

mp.dps = 50
C = 256


def companion(b, a):
    A = np.zeros((5, 5))
    for i in range(5):
        A[i, 0] = -a[i + 1]
        if i + 1 < 5:
            A[i, i + 1] = 1.0
    g = np.array([b[i + 1] - a[i + 1] * b[0] for i in range(5)])
    return A, g


def mp_power(A, n):
    M = matrix(A.tolist())
    return M ** n


def run_chunks(b, a, x, bounds, z0, states, A):
    """bounds: chunk boundaries over x (processing order); states[c]: zero-state final state of chunk c in
    float64.  Carry in mp, final pass = scipy's recurrence from the float64-rounded entry state."""
    y = np.empty_like(x)
    z = matrix([mpf(float(v)) for v in z0])
    powers = {}
    for c in range(len(bounds) - 1):
        lo, hi = bounds[c], bounds[c + 1]
        zin = np.array([float(v) for v in z])
        y[lo:hi], _ = signal.lfilter(b, a, x[lo:hi], zi=zin)
        n = hi - lo
        if n not in powers:
            powers[n] = mp_power(A, n)
        z = powers[n] * z + matrix([mpf(float(v)) for v in states[c]])
    return y


def main():
    b, a = signal.butter(5, 30 / 8000.0, btype="high")
    zi = signal.lfilter_zi(b, a)
    A, g = companion(b, a)
    # coefficient table of the dot products: c[j] = A^j g, j < 256 (formed in mp, rounded to float64)
    coef = np.zeros((C, 5))
    v = matrix([mpf(float(t)) for t in g])
    M = matrix(A.tolist())
    for j in range(C):
        coef[j] = [float(t) for t in v]
        v = M * v
    print("max |A^j g| over j < 256: %.4f   (max |A^256| entry: %.3e)" % (np.abs(coef).max(), max(abs(t) for t in mp_power(A, C))))

    rng = np.random.default_rng(0)
    for L, kind in ((48000, "speech-like"), (48128, "speech-like, L % 256 == 0"), (20011, "low-frequency heavy")):
        t = np.arange(L) / 16000.0
        if kind.startswith("low"):
            x = 0.35 * np.sin(2 * np.pi * 40 * t) + 0.1 * np.sin(2 * np.pi * 7 * t) + 1e-3 * rng.standard_normal(L)
        else:
            x = 0.3 * np.sin(2 * np.pi * 120 * t) * (np.sin(2 * np.pi * 1.1 * t) > -0.2) + 0.02 * rng.standard_normal(L)
        if L % 256 == 0:
            x = np.concatenate([x, [1e-6]])
        ext = np.concatenate([2 * x[0] - x[18:0:-1], x, 2 * x[-1] - x[-2:-20:-1]])
        y1, _ = signal.lfilter(b, a, ext, zi=zi * ext[0])
        r = y1[::-1].copy()
        N = r.shape[0]
        z0 = zi * r[0]
        ref, _ = signal.lfilter(b, a, r, zi=z0)

        # scheme A: chunks from the start of the reversed signal, zero-state recurrence for the states
        bA = list(range(0, N, C)) + [N]
        sA = [signal.lfilter(b, a, r[lo:hi], zi=np.zeros(5))[1] for lo, hi in zip(bA[:-1], bA[1:])]
        yA = run_chunks(b, a, r, bA, z0, sA, A)

        # scheme B: forward-aligned chunks.  Forward chunk k covers y1[256 k, 256 k + 256); in the reversed
        # signal that is r[N - 256 k - 256, N - 256 k).  The first backward chunk is the partial one.
        first = N % C if N % C else C
        bB = [0] + list(range(first, N, C)) + ([N] if (N - first) % C else [])
        if bB[-1] != N:
            bB.append(N)
        sB = []
        for lo, hi in zip(bB[:-1], bB[1:]):
            seg_fwd = r[lo:hi][::-1]                       # the chunk's y1 values in FORWARD order
            n = hi - lo
            # last processed sample (forward index 0 of the chunk) has coefficient A^0 g
            sB.append(sum(coef[j] * seg_fwd[j] for j in range(n)))          # forward-order accumulation
        yB = run_chunks(b, a, r, bB, z0, sB, A)
        print("L = %6d (%s): max|A - ref| = %.3e   max|B - ref| = %.3e   max|ref| = %.3f"
              % (L, kind, np.abs(yA - ref).max(), np.abs(yB - ref).max(), np.abs(ref).max()))


if __name__ == "__main__":
    main()
