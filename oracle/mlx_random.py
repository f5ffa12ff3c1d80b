"""Restatement of MLX's counter-based RNG -- TEST INFRASTRUCTURE ONLY.

Needed for one thing: the reference's only value-level test on this path,
`test_rope` (mlx-rs/src/fast.rs:231-251, mlx-rs/src/nn/positional_encoding.rs:
432-463), draws its input from `random::seed(71)` + `uniform(0, 1, [2, 8, 16])`.
To replay that golden vector without MLX the key schedule has to be restated:

* `RandomState::seed` / `next` (mlx-rs/src/random.rs:27-43): state = key(seed);
  every draw does `(state, subkey) = split(state, 2)` and uses `subkey`.
* `key`, `split`, `bits`, `uniform`: ml-explore/mlx v0.30.1 `mlx/random.cpp`
  and the CPU `RandomBits` kernel (threefry2x32, 20 rounds, the JAX layout:
  counters (i, i + half) fill outputs i and i + half).  Not vendored in
  /root/reference; restated from the published algorithm and VERIFIED by the
  input statistics the reference test itself asserts (mean 0.5082664489746094,
  sum 130.1162109375) -- see tests/test_oracle_golden.py.
"""
import numpy as np

_M = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) & _M) | (x >> (32 - r))


def threefry2x32(key, count):
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    ks = (key[0], key[1], key[0] ^ key[1] ^ 0x1BD11BDA)
    a = (count[0] + ks[0]) & _M
    b = (count[1] + ks[1]) & _M
    for i in range(5):
        for r in rot[i % 2]:
            a = (a + b) & _M
            b = _rotl(b, r)
            b ^= a
        a = (a + ks[(i + 1) % 3]) & _M
        b = (b + ks[(i + 2) % 3] + i + 1) & _M
    return a, b


def bits_u32(key, n):
    """n uint32 words from one key (RandomBits CPU kernel, width 4)."""
    out = [0] * n
    half = n // 2
    even = n % 2 == 0
    c0, c1 = 0, half + (0 if even else 1)
    while c0 + 1 < half:
        out[c0], out[c1] = threefry2x32(key, (c0, c1))
        c0 += 1
        c1 += 1
    if c0 < half:
        out[c0], out[c1] = threefry2x32(key, (c0, c1))
        c0 += 1
    if not even:
        out[half] = threefry2x32(key, (c0, 0))[0]
    return out


def key(seed):
    return ((seed >> 32) & _M, seed & _M)


def split(k):
    o = bits_u32(k, 4)
    return (o[0], o[1]), (o[2], o[3])


class RandomState:
    """mlx-rs/src/random.rs:20-43."""

    def __init__(self, seed):
        self.state = key(seed)

    def next(self):
        self.state, sub = split(self.state)
        return sub


def uniform_f32(state, shape, low=0.0, high=1.0):
    """mlx::core::random::uniform for float32."""
    n = int(np.prod(shape))
    u = np.array(bits_u32(state.next(), n), dtype=np.uint32)
    x = u.astype(np.float32) / np.float32(4294967295.0)
    x = np.minimum(x, np.nextafter(np.float32(1), np.float32(0)))
    rng = np.float32(high) - np.float32(low)
    return (rng * x + np.float32(low)).astype(np.float32).reshape(shape)
