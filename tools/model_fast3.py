"""numpy model of the three-pass register kernel (fast3): validates the index formulas, smem layouts,
twiddle tables and thread ownership before they are transcribed to CUDA."""
import numpy as np

def model(N, R1, R2, R3, E, x, check_conflicts=True):
    assert R1 * R2 * R3 == N and N % E == 0
    T = N // E
    assert E % R1 == 0 and E % R2 == 0 and E % R3 == 0 and T % R1 == 0
    M1 = N // R1           # = R2*R3
    W = np.exp(-2j * np.pi * np.arange(N) / N)
    # ---- load: regs[t][q] = x[t + T*q]
    regs = np.array([[x[t + T * q] for q in range(E)] for t in range(T)])
    # ---- pass 1
    P1 = ((M1 + 7) // 8) * 8 + 1
    X1 = np.zeros(R1 * P1, complex)
    for t in range(T):
        for m in range(E // R1):
            i1 = t + T * m
            v = np.array([regs[t][m + (E // R1) * j1] for j1 in range(R1)])
            y = np.fft.fft(v)
            for k1 in range(R1):
                X1[k1 * P1 + i1] = y[k1] * W[(i1 * k1) % N]          # tw1[k1][i1]
    # ---- pass 2: butterflies b2 = t + T*m2: k1 = b2 % R1, i2 = b2 // R1
    P2 = R1 * R2
    X2 = np.zeros(R3 * P2, complex)
    conflicts = 0
    for m2 in range(E // R2):
        for j2 in range(R2):
            addrs = []
            for t in range(T):
                b2 = t + T * m2; k1 = b2 % R1; i2 = b2 // R1
                addrs.append(k1 * P1 + i2 + R3 * j2)
            if check_conflicts:
                for q0 in range(0, T, 8):
                    s = [a % 8 for a in addrs[q0:q0 + 8]]
                    conflicts += len(s) - len(set(s))
    for t in range(T):
        for m2 in range(E // R2):
            b2 = t + T * m2; k1 = b2 % R1; i2 = b2 // R1
            v = np.array([X1[k1 * P1 + i2 + R3 * j2] for j2 in range(R2)])
            y = np.fft.fft(v)
            for k2 in range(R2):
                X2[i2 * P2 + k1 + R1 * k2] = y[k2] * W[(R1 * i2 * k2) % N]   # tw2[k2][i2]
    # ---- pass 3: butterflies klow = t + T*m3
    out = np.zeros(N, complex)
    for t in range(T):
        for m3 in range(E // R3):
            klow = t + T * m3
            v = np.array([X2[j3 * P2 + klow] for j3 in range(R3)])
            y = np.fft.fft(v)
            for k3 in range(R3):
                out[klow + R1 * R2 * k3] = y[k3]
    return out, conflicts

rng = np.random.default_rng(0)
for (N, R1, R2, R3, E) in ((4096, 16, 16, 16, 16), (2048, 8, 16, 16, 16), (2048, 16, 16, 8, 16), (8192, 32, 16, 16, 32), (8192, 16, 16, 32, 32),
                           (512, 8, 8, 8, 8), (1000, 10, 10, 10, 10), (500, 5, 10, 10, 10)):
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    try:
        out, conf = model(N, R1, R2, R3, E, x)
        print(N, (R1, R2, R3), E, "T=", N // E, "err", np.abs(out - np.fft.fft(x)).max(), "pass2-read conflicts", conf)
    except AssertionError as e:
        print(N, (R1, R2, R3), E, "constraint violated")


def model_r2c_pair(N, R1, R2, R3, E, xr):
    """r2c of 2N reals with the Hermitian post-twiddle done INSIDE pass 3: a thread runs the pass-3 butterflies
    klow = u and P2 - u, whose outputs are each other's mirrors (bin k = u + P2*k3 <-> N - k = (P2-u) + P2*(R3-1-k3)),
    so Z[k] and Z[N-k] meet in registers.  Unit 0 = butterflies 0 and P2/2 (each mirrors into itself)."""
    T = N // E
    P2 = R1 * R2
    assert P2 % 2 == 0
    z = xr[0::2] + 1j * xr[1::2]
    Z, _ = model(N, R1, R2, R3, E, z, check_conflicts=False)     # Z[klow + P2*k3] is what butterfly klow leaves in y[k3]
    out = np.full(N + 1, np.nan, complex)
    units = P2 // 2
    NU = -(-units // T)                                           # units per thread (last sweep may be ragged)
    w2n = lambda k: np.exp(-1j * np.pi * k / N)
    owner = {}
    for t in range(T):
        for mu in range(NU):
            u = t + T * mu
            if u >= units:
                continue
            if u == 0:
                ya = np.array([Z[0 + P2 * k3] for k3 in range(R3)]); yb = np.array([Z[P2 // 2 + P2 * k3] for k3 in range(R3)])
                out[0] = ya[0].real + ya[0].imag; out[N] = ya[0].real - ya[0].imag
                for k3 in range(1, R3 // 2 + 1):
                    a, c = ya[k3], ya[R3 - k3]; k = P2 * k3
                    s, d = a + np.conj(c), a - np.conj(c)
                    Q = 0.5j * w2n(k) * d
                    out[k] = 0.5 * s - Q; out[N - k] = np.conj(0.5 * s + Q)
                for k3 in range((R3 + 1) // 2):
                    a, c = yb[k3], yb[R3 - 1 - k3]; k = P2 // 2 + P2 * k3
                    s, d = a + np.conj(c), a - np.conj(c)
                    Q = 0.5j * w2n(k) * d
                    out[k] = 0.5 * s - Q; out[N - k] = np.conj(0.5 * s + Q)
                continue
            ka, kb = u, P2 - u
            ya = np.array([Z[ka + P2 * k3] for k3 in range(R3)]); yb = np.array([Z[kb + P2 * k3] for k3 in range(R3)])
            wt = w2n(u)                                           # per-thread factor, twr[u]
            for k3 in range(R3):
                a, c = ya[k3], yb[R3 - 1 - k3]; k = ka + P2 * k3
                root = np.exp(-2j * np.pi * k3 / (2 * R3))        # compile-time root W_{2 R3}^{k3}
                s, d = a + np.conj(c), a - np.conj(c)
                Q = 0.5j * (wt * root) * d
                assert k not in owner and (N - k) not in owner
                owner[k] = owner[N - k] = t
                out[k] = 0.5 * s - Q; out[N - k] = np.conj(0.5 * s + Q)
    return out


for (N, R1, R2, R3, E) in ((2048, 16, 16, 8, 16), (1024, 16, 8, 8, 16), (256, 8, 8, 4, 8), (500, 10, 10, 5, 10), (1944, 18, 18, 6, 18), (1000, 10, 10, 10, 10), (4096, 16, 16, 16, 16)):
    xr = rng.standard_normal(2 * N)
    out = model_r2c_pair(N, R1, R2, R3, E, xr)
    print("r2c pair", 2 * N, (R1, R2, R3), E, "err", np.abs(out - np.fft.rfft(xr)).max())
