"""Host-side threefry key bookkeeping for the gym adapters (jax.random.PRNGKey / split semantics; the
per-env keys themselves are derived on the device by pobrax_split_keys). Pure Python integers."""
from typing import Tuple

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def prng_key(seed: int) -> Tuple[int, int]:
    """jax.random.PRNGKey(seed) = (hi32, lo32)."""
    seed &= 0xFFFFFFFFFFFFFFFF
    return (seed >> 32, seed & _M)


def threefry2x32(key: Tuple[int, int], x0: int, x1: int) -> Tuple[int, int]:
    ks = (key[0], key[1], key[0] ^ key[1] ^ 0x1BD11BDA)
    x0, x1 = (x0 + ks[0]) & _M, (x1 + ks[1]) & _M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def split_at(key: Tuple[int, int], num: int, j: int) -> Tuple[int, int]:
    """jax.random.split(key, num)[j] (non-partitionable threefry: flat = random_bits(key, 2*num))."""
    out = []
    for i in (2 * j, 2 * j + 1):
        a = i if i < num else i - num
        y0, y1 = threefry2x32(key, a, a + num)
        out.append(y0 if i < num else y1)
    return out[0], out[1]
