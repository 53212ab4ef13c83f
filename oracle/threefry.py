"""CPU restatement of jax.random (threefry2x32, non-partitionable scheme, jax 0.3-era).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product path (po_brax_b200/) never imports it.

The reference reaches this code through un-vendored third-party packages:
  * brax.jumpy.random_split / random_uniform / random_prngkey (dispatch to jax.random when
    the key is a tracer, i.e. under every vmap/jit path: /root/reference/po_brax/envs/wrappers.py:13,166,172)
  * jax.random.randint  (/root/reference/po_brax/envs/ant_tag.py:132)
  * jax.random.choice   (/root/reference/po_brax/more_jp.py:75)
jax is not installed here and is unpinned by the reference (setup.py:14 pins only brax>=0.0.12);
this file restates the published algorithm (Salmon et al. Threefry-2x32-20 as used by jax._src.prng:
threefry_2x32, threefry_split, threefry_random_bits; jax._src.random: uniform, randint, _shuffle,
choice) and is pinned by the public known-answer vectors in tests/test_threefry.py
(Random123 KATs + jax's documented split(PRNGKey(0)) / uniform(PRNGKey(0))).

All functions are batched over leading key dimensions: key[..., 2] uint32.
"""
import numpy as np

U32 = np.uint32
_R0 = (13, 15, 26, 6)
_R1 = (17, 29, 16, 24)


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) = (hi32, lo32)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """One Threefry-2x32-20 block per element. All args uint32 arrays (broadcastable)."""
    with np.errstate(over='ignore'):
        k0 = np.asarray(k0, U32)
        k1 = np.asarray(k1, U32)
        x0 = np.asarray(x0, U32).copy()
        x1 = np.asarray(x1, U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in (_R0 if i % 2 == 0 else _R1):
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def random_bits(key, n: int) -> np.ndarray:
    """threefry_random_bits(key, 32, (n,)): key[..., 2] -> bits[..., n]."""
    key = np.asarray(key, U32)
    m = n + (n & 1)
    cnt = np.arange(m, dtype=U32)
    cnt[n:] = 0  # odd n: padded with a single 0
    h = m // 2
    y0, y1 = threefry2x32(key[..., 0:1], key[..., 1:2], cnt[:h], cnt[h:])
    return np.concatenate([y0, y1], axis=-1)[..., :n]


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split: key[..., 2] -> keys[..., num, 2]."""
    bits = random_bits(key, 2 * num)
    return bits.reshape(bits.shape[:-1] + (num, 2))


def uniform(key, n: int, lo, hi) -> np.ndarray:
    """jax.random.uniform(key, (n,), float32, lo, hi); lo/hi scalars or [n] arrays."""
    bits = random_bits(key, n)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def randint(key, lo: int, hi: int) -> np.ndarray:
    """jax.random.randint(key, (), lo, hi) for int32; returns int32[...]."""
    ks = split(key, 2)
    hi_bits = random_bits(ks[..., 0, :], 1)[..., 0]
    lo_bits = random_bits(ks[..., 1, :], 1)[..., 0]
    span = U32(hi - lo) if hi > lo else U32(1)
    mult = U32(2 ** 16) % span
    with np.errstate(over='ignore'):
        mult = (mult * mult) % span
        off = ((hi_bits % span) * mult + (lo_bits % span)) % span
    return (off.astype(np.int64) + lo).astype(np.int32)


def shuffle_indices(key, n: int) -> np.ndarray:
    """jax.random._shuffle(key, arange(n)): rounds of stable sort by fresh random u32 keys."""
    key = np.asarray(key, U32)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    idx = np.broadcast_to(np.arange(n, dtype=np.int32), key.shape[:-1] + (n,)).copy()
    for _ in range(rounds):
        ks = split(key, 2)
        key, sub = ks[..., 0, :], ks[..., 1, :]
        sort_keys = random_bits(sub, n)
        order = np.argsort(sort_keys, axis=-1, kind='stable')
        idx = np.take_along_axis(idx, order, axis=-1)
    return idx


def choice_no_replace(key, n: int, n_draws: int) -> np.ndarray:
    """Indices drawn by jax.random.choice(key, a[n, ...], (n_draws,), replace=False, axis=0)."""
    return shuffle_indices(key, n)[..., :n_draws]
