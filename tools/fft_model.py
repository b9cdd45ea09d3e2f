"""
Design aid (not product code): NumPy model of the register-resident Stockham
FFT used by syncopy_b200/csrc/fft_core.cuh, plus a shared-memory bank-conflict
counter for the exchange index patterns.  Run: python tools/fft_model.py
"""
import numpy as np


def radices_for(log2n, emax=16):
    """As many radix-`emax` passes as possible, remainder first."""
    le = int(np.log2(emax))
    q, rem = divmod(log2n, le)
    r = ([1 << rem] if rem else []) + [emax] * q
    return r


def stockham(x, radices, E):
    N = x.size
    NT = N // E
    v = np.zeros((NT, E), complex)          # thread t holds index t + NT*e
    for t in range(NT):
        v[t] = x[t + NT * np.arange(E)]
    Ns = 1
    for R in radices:
        smem = np.zeros(N, complex)
        for t in range(NT):
            for u in range(E // R):
                b = t + u * NT
                k = b % Ns
                xin = np.array([v[t, u + r * (E // R)] for r in range(R)])
                tw = np.exp(-2j * np.pi * k * np.arange(R) / (Ns * R))
                xin = xin * tw
                y = np.fft.fft(xin)
                j0 = (b // Ns) * Ns * R + k
                for q in range(R):
                    smem[j0 + q * Ns] = y[q]
        for t in range(NT):
            v[t] = smem[t + NT * np.arange(E)]
        Ns *= R
    return smem


def conflicts(addr_words, word_bytes=8):
    """Number of shared-memory wavefronts for one warp access of `word_bytes`-wide words."""
    lanes_per_phase = 32 * 4 // word_bytes // 4 * 4 if word_bytes > 4 else 32
    lanes_per_phase = {4: 32, 8: 16, 16: 8}[word_bytes]
    total = 0
    for ph in range(0, 32, lanes_per_phase):
        a = addr_words[ph:ph + lanes_per_phase]
        banks = {}
        for w in a:
            base = (w * word_bytes // 4) % 32
            banks.setdefault(base, set()).add(w)
        total += max(len(s) for s in banks.values())
    return total


def exchange_conflicts(N, radices, E, P, pad):
    """
    Thread mapping: tid = j*P + p (p = channel-pair lane, j = thread of the pair FFT).
    smem word address (float2 units) = pad(idx)*P + p.
    """
    NT = N // E
    res = []
    Ns = 1
    for pi, R in enumerate(radices):
        worst_w = worst_r = 0
        for warp in range(max(1, NT * P // 32)):
            tids = np.arange(32) + 32 * warp
            j, p = tids // P, tids % P
            j = j % NT
            for u in range(E // R):
                b = j + u * NT
                k = b % Ns
                j0 = (b // Ns) * Ns * R + k
                for q in range(R):
                    idx = j0 + q * Ns
                    worst_w = max(worst_w, conflicts(list(pad(idx) * P + p)))
            for e in range(E):
                idx = j + NT * e
                worst_r = max(worst_r, conflicts(list(pad(idx) * P + p)))
        res.append((R, Ns, worst_w, worst_r))
        Ns *= R
    return res


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for L in range(4, 13):
        N = 1 << L
        E = min(16, N)
        rad = radices_for(L)
        x = rng.normal(size=N) + 1j * rng.normal(size=N)
        y = stockham(x, rad, E)
        err = np.abs(y - np.fft.fft(x)).max()
        print(f"N={N:5d} radices={rad} err={err:.1e}")
    print("bank conflicts (ideal = 2 wavefronts for 8-byte words per warp):")
    for L in (6, 8, 9, 10, 11, 12, 13, 14):
        N = 1 << L
        rad = radices_for(L)
        for P in (4,):
            for name, pad in (("none", lambda i: i), ("i+i/16", lambda i: i + (i >> 4)),
                              ("i+i/32", lambda i: i + (i >> 5))):
                print(f"  N={N} P={P} pad={name}: ", exchange_conflicts(N, rad, 16, P, pad))


def radices_small_last(log2n, emax=16):
    le = int(np.log2(emax))
    q, rem = divmod(log2n, le)
    return [emax] * q + ([1 << rem] if rem else [])
