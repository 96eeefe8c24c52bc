"""Model of shared-memory bank conflicts for the mixed-radix exchange layout used by
csrc/thz_fft.cuh (E=16 elements/thread, padded float2 array: addr(i) = i + (i >> PAD_SHIFT)).
Run: python tools/bank_conflicts.py"""
import itertools

PLANS = {64: [16, 4], 128: [16, 8], 256: [16, 16], 512: [16, 8, 4], 1024: [16, 16, 4],
         2048: [16, 16, 8], 4096: [16, 16, 16], 8192: [16, 16, 8, 4]}
E = 16


def pad(i, shift=4):
    return i + (i >> shift)


def stage_indices(N, radices, s, t):
    """element indices (in register order u + U*m) thread t owns at stage s"""
    T = N // E
    L = N
    for r in radices[:s]:
        L //= r
    R = radices[s]
    S = L // R
    U = E // R
    idx = [0] * E
    for u in range(U):
        beta = t + u * T
        b, j = divmod(beta, S)
        for m in range(R):
            idx[u + U * m] = b * L + j + m * S
    return idx


def natural_indices(N, radices, t):
    """natural-order bin k held in register (u,m) after the last DIF stage"""
    pos = stage_indices(N, radices, len(radices) - 1, t)
    return [pos_to_k(N, radices, p) for p in pos]


def pos_to_k(N, radices, p):
    # position p = q0*S0 + q1*S1 + ... ; k = q0 + R0*q1 + R0*R1*q2 ...
    L = N
    k = 0
    w = 1
    for r in radices:
        S = L // r
        q, p = divmod(p, S)
        k += q * w
        w *= r
        L = S
    return k


def conflict_degree(addrs_f2):
    """addrs_f2: 16 float2 addresses of one half-warp -> max words per bank"""
    banks = {}
    for a in set(addrs_f2):
        for w in (2 * a, 2 * a + 1):
            banks.setdefault(w % 32, set()).add(w)
    return max(len(v) for v in banks.values())


def check(N, radices, shift=4):
    T = N // E
    nthreads = max(T, 32)
    res = []
    for s in range(len(radices)):
        worst = 0
        for h0 in range(0, nthreads, 16):
            lanes = [t % T for t in range(h0, h0 + 16)]
            grp = [(t // T) for t in range(h0, h0 + 16)]
            for reg in range(E):
                addrs = [pad(stage_indices(N, radices, s, lanes[l])[reg], shift) + grp[l] * pad(N, shift)
                         for l in range(16)]
                worst = max(worst, conflict_degree(addrs))
        res.append(worst)
    # natural-order scatter after last stage
    worst = 0
    for h0 in range(0, nthreads, 16):
        lanes = [t % T for t in range(h0, h0 + 16)]
        grp = [(t // T) for t in range(h0, h0 + 16)]
        for reg in range(E):
            addrs = [pad(natural_indices(N, radices, lanes[l])[reg], shift) + grp[l] * pad(N, shift) for l in range(16)]
            worst = max(worst, conflict_degree(addrs))
    return res, worst


if __name__ == "__main__":
    for N, r in PLANS.items():
        print(N, r, check(N, r))
