"""philox.py -- TEST INFRASTRUCTURE: numpy restatement of the counter-based random inputs of csrc/rng.cu.

Philox4x32-10 is the published algorithm of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11; Random123 philox.h); pinned in tests/test_rng.py by the Random123 known-answer vectors.  The mapping
(walker, step, slot) -> numbers is the one documented in csrc/rng.cu: counter = (walker lo, walker hi, step, slot),
key = the 64-bit seed.  Uniform streams are bit-exact with the CUDA kernels; normals agree to libm rounding (log /
sincospi differ in the last ulp between glibc and the CUDA math library).  The reference itself draws from jax's
threefry (VMCmcstep.py:19-20,58,83; pseudopotential.py:234), which cannot be reproduced without jax -- RNG streams stay
"parity unpinned" against the reference; only the distributions and the stream's independence of sharding are claimed."""
import numpy as np

M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
SLOT_SWEEP, SLOT_ROT, SLOT_UNIFORM = 0, 0x10000, 0x20000


def philox4x32_10(ctr, key):
    """ctr (..., 4) uint32, key (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., k], dtype=np.uint32) for k in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32)
    k1 = np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            c = [(p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0, p1.astype(np.uint32),
                 (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1, p0.astype(np.uint32)]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def _u53(hi, lo):
    return ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26) | (lo.astype(np.uint64) >> np.uint64(6))).astype(np.float64) / 9007199254740992.0


def _draw(seed, walkers, step, slot):
    """walkers (...,) int64, slot (...,) -> (..., 4) uint32 words."""
    w = np.asarray(walkers, dtype=np.uint64)
    slot = np.broadcast_to(np.asarray(slot, dtype=np.uint32), w.shape)
    ctr = np.stack([(w & np.uint64(0xFFFFFFFF)).astype(np.uint32), (w >> np.uint64(32)).astype(np.uint32),
                    np.full(w.shape, step, dtype=np.uint32), slot], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def _normal_pair(r):
    u1 = 1.0 - _u53(r[..., 0], r[..., 1])
    u2 = _u53(r[..., 2], r[..., 3])
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)


def rng_sweep(seed, step, walker0, B, n, tstep):
    """-> gauss1 (B,3N), gauss2c (B,N,3), rnd (B,N): the arrays aiqmc_rng_sweep writes."""
    w = (walker0 + np.arange(B, dtype=np.int64))[:, None] + np.zeros((1, n), dtype=np.int64)
    i = np.arange(n, dtype=np.uint32)[None, :] + np.zeros((B, 1), dtype=np.uint32)
    v = []
    for c in range(3):
        a, b = _normal_pair(_draw(seed, w, step, SLOT_SWEEP + 4 * i + c))
        v += [a, b]
    r = _draw(seed, w, step, SLOT_SWEEP + 4 * i + 3)
    s = np.sqrt(tstep)
    gauss1 = s * np.stack(v[:3], axis=-1).reshape(B, 3 * n)
    gauss2c = s * np.stack(v[3:], axis=-1)
    return gauss1, gauss2c, _u53(r[..., 0], r[..., 1])


def rng_rotations(seed, step, walker0, B):
    w = walker0 + np.arange(B, dtype=np.int64)
    g = []
    for c in range(5):
        a, b = _normal_pair(_draw(seed, w, step, SLOT_ROT + c))
        g += [a, b]
    m = np.stack(g[:9], axis=-1).reshape(B, 3, 3)
    q, r = np.linalg.qr(m)
    return q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[:, None, :]


def rng_uniform(seed, step, walker0, B, cols, tag):
    w = (walker0 + np.arange(B, dtype=np.int64))[:, None] + np.zeros((1, cols), dtype=np.int64)
    c = np.arange(cols, dtype=np.uint32)[None, :] + np.zeros((B, 1), dtype=np.uint32)
    r = _draw(seed, w, step, SLOT_UNIFORM + tag * 0x1000 + c)
    return _u53(r[..., 0], r[..., 1])


def expand_gauss2(gauss2c):
    """(B,N,3) diagonal blocks -> the reference-shaped (B,N,3N) array with zeros elsewhere (only the diagonal blocks
    are ever read, VMCmcstep.py:86-94)."""
    B, n, _ = gauss2c.shape
    out = np.zeros((B, n, 3 * n))
    for i in range(n):
        out[:, i, 3 * i:3 * i + 3] = gauss2c[:, i]
    return out
