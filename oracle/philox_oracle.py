"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the engine's counter-based noise stream
(motion-diffusion-model_b200/csrc/kernels.cuh: philox_normal_kernel; include/b200mdm.h: b200mdm_philox_normal).

Not a restatement of anything in the reference: the reference draws `th.randn_like(x)` from torch's global generator
(diffusion/gaussian_diffusion.py:525, :691, :770).  The engine offers this stream as an alternative whose values depend
only on (seed, schedule index of the consuming step, GLOBAL sample index, element) so that sharding a batch over GPUs
cannot change a sample's noise.  Pinned by the Philox4x32-10 known-answer vectors of the Random123 distribution
(kat_vectors: counter/key all-zero, all-ones, and the pi digits case) -- see tests/test_philox_cpu.py.

    counter = (q, step_id, g_lo, g_hi ^ 0x4d444d42),  key = (seed_lo, seed_hi),  q = element // 4
    u_k = ((w_k >> 8) + 0.5) * 2^-24
    elements 4q..4q+3 = r0 cos(2 pi u1), r0 sin(2 pi u1), r1 cos(2 pi u3), r1 sin(2 pi u3);  r = sqrt(-2 ln u)
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint32 arrays (or scalars); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def normal(batch, n_per_sample, seed, sample_index_base, step_id):
    """fp32 [batch, n_per_sample]: what b200mdm_philox_normal writes (float32 arithmetic mirrored with numpy)."""
    qn = (n_per_sample + 3) // 4
    q = np.arange(qn, dtype=np.uint64)
    out = np.empty((batch, qn * 4), dtype=np.float32)
    seed = int(seed) & (2 ** 64 - 1)
    for b in range(batch):
        g = (int(sample_index_base) + b) & (2 ** 64 - 1)
        w = philox4x32_10(q, np.uint64(int(step_id) & 0xFFFFFFFF), np.uint64(g & 0xFFFFFFFF),
                          np.uint64(((g >> 32) ^ 0x4d444d42) & 0xFFFFFFFF), seed & 0xFFFFFFFF, seed >> 32)
        u = [((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24) for x in w]
        for h in range(2):
            r = np.sqrt(np.float32(-2.0) * np.log(u[2 * h])).astype(np.float32)
            ang = (np.float32(2.0) * u[2 * h + 1]).astype(np.float64) * np.pi
            out[b, 2 * h::4] = (r * np.cos(ang).astype(np.float32)).astype(np.float32)
            out[b, 2 * h + 1::4] = (r * np.sin(ang).astype(np.float32)).astype(np.float32)
    return out[:, :n_per_sample]
