"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the part of JAX's PRNG that the reference's tests use to create inputs, so that
reference tests whose expected values depend on ``jax.random`` can be replayed without JAX:
``jax.random.PRNGKey(seed)`` and ``jax.random.normal(key, shape, float32)``.

JAX is a third-party dependency that is not under /root/reference (pinned ``jax 0.10.1`` in the
reference's ``uv.lock``); the algorithm is restated from its published implementation
(``jax/_src/prng.py``: ``threefry_seed``, ``threefry2x32``, ``_threefry_random_bits_partitionable``;
``jax/_src/random.py``: ``_uniform``, ``_normal_real``):

  * key            = (seed >> 32, seed & 0xffffffff) as two uint32 words;
  * random bits    = with ``jax_threefry_partitionable`` (the default since JAX 0.5): element i of a
                     flat array gets ``b1 ^ b2`` where ``(b1, b2) = threefry2x32(key, (hi(i), lo(i)))``;
  * uniform [0,1)  = the top 23 bits as the mantissa of a float in [1, 2), minus 1;
  * normal         = sqrt(2) * erfinv(u) with u = max(lo, u01 * (1 - lo) + lo), lo = nextafter(-1, 0).

Pinned by (1) the Threefry-2x32 known-answer vector of the Random123 distribution and (2) the
reference's own test ``tests/goose/test_interface_liesel.py:12-44``: with
``x = normal(PRNGKey(1337), (500,))`` and ``y ~ Normal(mu, 1)``, ``log_prob`` is -719.46875 at
``mu = 0`` and -25616.779296875 at ``mu = 10`` — reproduced by this restatement
(tests/test_oracle_pins.py); the pre-0.5 bit layout gives -722.71 instead.
"""
import numpy as np
from scipy.special import erfinv

_U = np.uint32


def _rotl(x, r):
    return ((x << _U(r)) | (x >> _U(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds, on uint32 arrays x0, x1 with the key words k0, k1."""
    x0 = np.asarray(x0, dtype=np.uint32).copy()
    x1 = np.asarray(x1, dtype=np.uint32).copy()
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    ks = (_U(k0), _U(k1), _U(k0) ^ _U(k1) ^ _U(0x1BD11BDA))
    with np.errstate(over="ignore"):
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in rot[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r) ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + _U(i + 1)).astype(np.uint32)
    return x0, x1


def prng_key(seed: int):
    return (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF


def random_bits(key, n: int) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint64)
    hi, lo = (idx >> np.uint64(32)).astype(np.uint32), (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    b1, b2 = threefry2x32(key[0], key[1], hi, lo)
    return b1 ^ b2


def normal(key, n: int) -> np.ndarray:
    """``jax.random.normal(key, (n,), float32)``; erfinv is evaluated in float64 and rounded."""
    bits = random_bits(key, n)
    u01 = ((bits >> _U(9)) | np.float32(1.0).view(np.uint32)).view(np.float32) - np.float32(1.0)
    lo, hi = np.nextafter(np.float32(-1.0), np.float32(0.0)), np.float32(1.0)
    u = np.maximum(lo, (u01 * (hi - lo) + lo).astype(np.float32))
    return (np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64)).astype(np.float32)).astype(np.float32)
