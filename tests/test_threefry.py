"""Pins oracle/threefry.py against public known-answer vectors:
Random123 Threefry-2x32-20 KATs and jax's documented split(PRNGKey(0)) / uniform(PRNGKey(0))."""
import numpy as np

from oracle import threefry as tf


def test_random123_kats():
    kats = [((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
            ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
            ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
    for key, ctr, out in kats:
        y0, y1 = tf.threefry2x32(np.uint32(key[0]), np.uint32(key[1]), np.uint32(ctr[0]), np.uint32(ctr[1]))
        assert (int(y0), int(y1)) == out


def test_jax_split_and_uniform_of_key0():
    k = tf.prng_key(0)
    assert tf.split(k, 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    u = tf.uniform(k, 1, 0., 1.)
    assert np.float32(u[0]) == np.float32(0.41845703)


def test_split_is_counter_based():
    k = tf.prng_key(42)
    full = tf.split(k, 129)
    assert full.shape == (129, 2)
    # batched keys give the same answer as one-by-one
    ks = np.stack([tf.prng_key(s) for s in range(5)])
    b = tf.split(ks, 4)
    for i in range(5):
        assert (b[i] == tf.split(ks[i], 4)).all()


def test_randint_span4_is_low_bits_and_3():
    ks = tf.split(tf.prng_key(7), 64)
    r = tf.randint(ks, 0, 4)
    lo = tf.random_bits(tf.split(ks, 2)[:, 1], 1)[:, 0]
    assert (r == (lo & 3).astype(np.int32)).all()
    assert set(r.tolist()) == {0, 1, 2, 3}


def test_shuffle_is_permutation_and_stable():
    ks = tf.split(tf.prng_key(3), 32)
    idx = tf.shuffle_indices(ks, 156)
    assert (np.sort(idx, axis=-1) == np.arange(156)).all()
    c = tf.choice_no_replace(ks, 156, 16)
    assert c.shape == (32, 16) and (c == idx[:, :16]).all()
