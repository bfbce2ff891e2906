"""Philox4x32-10 word stream + CPython `random` arithmetic on top of it (TEST INFRASTRUCTURE).

The planners consume randomness as a stream of 32-bit words, interpreted exactly the way
CPython's `random` module interprets MT19937 output (reference call sites:
src/graph.py:96-98 `random.random()/random.uniform()`, src/algorithm.py:813-816 `random.choice`).
Word `i` of stream `(seed, stream_id)` is

    philox4x32_10(counter=(i>>2 & 0xffffffff, i>>34, stream_id_lo, stream_id_hi),
                  key=(seed_lo, seed_hi))[i & 3]

The CUDA path (csrc/rng.cuh) and the C oracle (oracle/rrt_oracle.c) implement the same function;
this numpy version feeds the unmodified Python reference in oracle/refharness.py.
"""
import random as _random

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox_words(seed: int, stream_id: int, n_words: int, first_word: int = 0) -> np.ndarray:
    """Return words [first_word, first_word + n_words) of the stream as uint32."""
    lo_blk = first_word >> 2
    hi_blk = (first_word + n_words + 3) >> 2
    blk = np.arange(lo_blk, hi_blk, dtype=np.uint64)
    c0 = blk & MASK
    c1 = blk >> np.uint64(32)
    c2 = np.full_like(blk, stream_id & 0xFFFFFFFF)
    c3 = np.full_like(blk, (stream_id >> 32) & 0xFFFFFFFF)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    out = np.stack([c0, c1, c2, c3], axis=1).reshape(-1).astype(np.uint32)
    off = first_word - (lo_blk << 2)
    return out[off:off + n_words].copy()


class WordStreamRandom(_random.Random):
    """`random.Random` whose entropy is an explicit uint32 word list.

    Only `random()` and `getrandbits()` are overridden; `uniform`, `choice`, `_randbelow` are
    CPython's own, so the arithmetic on top of the words is the stdlib's, not a restatement.
    """

    def __init__(self, words):
        super().__init__(0)
        self._words = [int(w) for w in words]
        self.cursor = 0

    def _next(self) -> int:
        w = self._words[self.cursor]   # IndexError = stream exhausted (harness sizes it generously)
        self.cursor += 1
        return w

    def random(self) -> float:
        a = self._next() >> 5
        b = self._next() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)

    def getrandbits(self, k: int) -> int:
        if k <= 0:
            raise ValueError("number of bits must be greater than zero")
        if k <= 32:
            return self._next() >> (32 - k)
        # wider requests are never made on this path (lists stay < 2**32 long)
        raise NotImplementedError

    def seed(self, *a, **k):  # Random.__init__ calls seed(); keep it inert
        return None
