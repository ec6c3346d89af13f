"""CPU check of the fragment index maps of the tf32 mma.sync Dense kernels (no rounding: exact fp64)."""
import numpy as np
rng = np.random.default_rng(0)

def mma(c, a, b):
    """c[lane][4] += A(16x8) @ B(8x8) with per-lane fragments a[lane][4], b[lane][2] (m16n8k8 layouts)."""
    A = np.zeros((16, 8)); B = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
        B[t, g], B[t + 4, g] = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]

# ---- forward: y[n][o] = sum_k x[n][k] W[k][o]; warp: 32 columns (4 n-tiles), k-steps of 8
n, K, O = 5, 24, 40
x = rng.standard_normal((16, K)); x[n:] = 0
W = rng.standard_normal((K, O))
y = np.zeros((16, O))
for o0w in range(0, O, 32):
    for j in range(4):
        c = np.zeros((32, 4))
        for k in range(0, K, 8):
            a = np.zeros((32, 4)); b = np.zeros((32, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                a[lane] = [x[g, k + t], x[g + 8, k + t], x[g, k + t + 4], x[g + 8, k + t + 4]]
                o = o0w + 8 * j + g
                b[lane] = [W[k + t, o] if o < O else 0, W[k + t + 4, o] if o < O else 0]
            mma(c, a, b)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e, (r, cc) in enumerate([(g, 2 * t), (g, 2 * t + 1), (g + 8, 2 * t), (g + 8, 2 * t + 1)]):
                o = o0w + 8 * j + cc
                if o < O: y[r, o] = c[lane][e]
assert np.allclose(y[:n], (x @ W)[:n]), "fwd"

# ---- dgrad: dx[n][kin] = sum_o dy[n][o] W[kin][o]; warp: 8 kin rows; groups of 32 o, permuted k slots
K, O = 24, 64
dy = rng.standard_normal((16, O)); dy[n:] = 0
W = rng.standard_normal((K, O))
dx = np.zeros((16, K))
for kin0 in range(0, K, 8):
    c = np.zeros((32, 4))
    for grp in range(O // 32):
        for s in range(4):
            a = np.zeros((32, 4)); b = np.zeros((32, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                o = 32 * grp + 8 * t + 2 * s
                a[lane] = [dy[g, o], dy[g + 8, o], dy[g, o + 1], dy[g + 8, o + 1]]
                b[lane] = [W[kin0 + g, o], W[kin0 + g, o + 1]]
            mma(c, a, b)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        dx[g, kin0 + 2 * t], dx[g, kin0 + 2 * t + 1], dx[g + 8, kin0 + 2 * t], dx[g + 8, kin0 + 2 * t + 1] = c[lane]
assert np.allclose(dx[:n], (dy @ W.T)[:n]), "dgrad"

# ---- wgrad: dW[kin][o] = sum_n x[n][kin] dy[n][o]; warp: 16 kin rows, o tiles of 8, two k-steps over the batch
K, O = 32, 24
x = rng.standard_normal((16, K)); x[n:] = 0
dy = rng.standard_normal((16, O)); dy[n:] = 0
dW = np.zeros((K, O))
for kin0 in range(0, K, 16):
    for o0 in range(0, O, 8):
        c = np.zeros((32, 4))
        for ks in range(2):
            a = np.zeros((32, 4)); b = np.zeros((32, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                bt = 8 * ks + t
                a[lane] = [x[bt, kin0 + g], x[bt, kin0 + g + 8], x[bt + 4, kin0 + g], x[bt + 4, kin0 + g + 8]]
                b[lane] = [dy[bt, o0 + g], dy[bt + 4, o0 + g]]
            mma(c, a, b)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            dW[kin0 + g, o0 + 2 * t], dW[kin0 + g, o0 + 2 * t + 1], dW[kin0 + g + 8, o0 + 2 * t], dW[kin0 + g + 8, o0 + 2 * t + 1] = c[lane]
assert np.allclose(dW, x.T @ dy), "wgrad"
print("fragment maps OK")
