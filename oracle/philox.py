"""Philox4x32-10 (Random123) and the element-indexed noise streams of libed2b200, restated in NumPy.

Third-party algorithm: the reference draws noise through tf.random.* / tfp.distributions (un-vendored, un-pinned;
edward2/tensorflow/random_variable.py:161-174, edward2/tensorflow/layers/dense.py:261-270), whose streams are not
part of its contract.  This repo's native sampler is Philox4x32-10 as published in Salmon et al. (SC'11) /
Random123 1.14 `philox.h`; known-answer vectors from Random123 `kat_vectors` are in tests/golden/philox_kat.json.
Stream contract (must match edward2_b200/csrc/philox.cuh): key = (seed_lo, seed_hi),
counter = (block_lo, block_hi, stream_lo, stream_hi), block = element_index // 4.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: uint32 [..., 2] -> uint32 [..., 4]."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k0 = np.array(key[..., 0], dtype=np.uint32, copy=True)
    k1 = np.array(key[..., 1], dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def blocks(seed, stream, n_blocks, first_block=0):
    """uint32 [n_blocks, 4]: Philox output for blocks first_block .. first_block+n_blocks-1."""
    b = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    ctr = np.empty((n_blocks, 4), dtype=np.uint32)
    ctr[:, 0] = (b & MASK).astype(np.uint32)
    ctr[:, 1] = (b >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(seed_lo(stream))
    ctr[:, 3] = np.uint32(seed_hi(stream))
    key = np.empty((n_blocks, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed_lo(seed))
    key[:, 1] = np.uint32(seed_hi(seed))
    return philox4x32_10(ctr, key)


def seed_lo(x):
    return int(x) & 0xFFFFFFFF


def seed_hi(x):
    return (int(x) >> 32) & 0xFFFFFFFF


def _u32_to_uniform(x):
    return ((x >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24


def uniform(seed, stream, n):
    x = blocks(seed, stream, (n + 3) // 4).reshape(-1)[:n]
    return _u32_to_uniform(x)


def normal(seed, stream, n):
    """Eight normals per block (philox.cuh): word w of block i // 8 gives (z_2w, z_2w+1) = Box-Muller(u0, u1) with the
    16-bit uniforms u0 = (hi16 + 0.5) / 65536, u1 = (lo16 + 0.5) / 65536 (float64 evaluation of the fp32 recipe)."""
    x = blocks(seed, stream, (n + 7) // 8).reshape(-1)
    u0 = ((x >> np.uint32(16)).astype(np.float64) + 0.5) * 2.0 ** -16
    u1 = ((x & np.uint32(0xFFFF)).astype(np.float64) + 0.5) * 2.0 ** -16
    r = np.sqrt(-2.0 * np.log(u0))
    z = np.stack([r * np.cos(2 * np.pi * u1), r * np.sin(2 * np.pi * u1)], axis=-1)
    return z.reshape(-1)[:n]


def sign(seed, stream, n):
    """One bit per element: element i = bit (i % 32) of word (i // 32) % 4 of block i // 128 (philox.cuh)."""
    x = blocks(seed, stream, (n + 127) // 128).reshape(-1)  # uint32 words, 32 elements each
    bits = (x[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & np.uint32(1)
    return np.where(bits.reshape(-1)[:n] != 0, 1.0, -1.0)


def global_rows(n_local_rows, row0=0, rows_per_group_local=0, rows_per_group_global=0):
    """Row mapping of per-example noise tensors under data parallelism (ed2_noise in include/ed2b200.h,
    `noise_row` in edward2_b200/csrc/philox.cuh): global row of every local row."""
    m = np.arange(n_local_rows, dtype=np.int64)
    if rows_per_group_local > 0:
        g = m // rows_per_group_local
        return m + g * (rows_per_group_global - rows_per_group_local) + row0
    return m + row0


def per_example(kind, seed, stream, rows, cols, row0=0, rows_per_group_local=0, rows_per_group_global=0):
    """[rows, cols] per-example noise (kind 'normal' | 'sign') of a rank holding `rows` local rows: element (m, c) is
    element global_row(m) * cols + c of the stream, i.e. the rows of the single-device global-batch tensor."""
    gr = global_rows(rows, row0, rows_per_group_local, rows_per_group_global)
    total = int(gr.max() + 1) * cols if rows else 0
    flat = {"normal": normal, "sign": sign}[kind](seed, stream, total)
    return flat.reshape(-1, cols)[gr]
