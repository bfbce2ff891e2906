// rng.cuh — 32-bit random word source + CPython `random` arithmetic on top of it.
//
// The reference draws from Python's global `random` (src/graph.py:96-98, src/algorithm.py:813-816).
// Here the planner consumes a stream of 32-bit words and applies CPython's own formulas:
//   random()      = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53
//   uniform(a, b) = a + (b - a) * random()
//   choice(seq)   = seq[_randbelow(len)],  _randbelow(n): k = n.bit_length(); r = w >> (32 - k) until r < n
// The words come either from an explicit buffer (drop-in mode: the MT19937 words `random` itself would
// have produced) or from Philox4x32-10 keyed by (seed, stream_id).
#pragma once
#include <stdint.h>

#include "../../include/rrtfnd.h"

namespace rrtfnd {

struct WordSrc {
    const uint32_t* words;  // STREAM_WORDS buffer of this planner
    long long n_words;
    unsigned long long seed, stream, cursor, blk_idx;
    uint32_t b0, b1, b2, b3;   // cached Philox block (scalars: a dynamically indexed array would live in local memory)
    int mode;
    bool have_blk, exhausted;
};

// One Philox4x32-10 block. Not inlined: the ten rounds exist once in the kernel however many draws it makes.
static __device__ __noinline__ uint4 philox4x32_10(unsigned long long blk, unsigned long long stream, unsigned long long seed) {
    uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The word source of a planner lives in shared memory (one copy per warp, every lane reads and writes the same
// values), so drawing a word is an ordinary call instead of a block of code inlined at every draw.
// Every lane stores the same values; the __syncwarp()s keep the read-modify-write of the shared copy race free even if the
// warp reaches the call in diverged sub-groups (independent thread scheduling): all lanes read, then all lanes write.
static __device__ __noinline__ uint32_t next_word(WordSrc& r) {
    __syncwarp();
    const unsigned long long i = r.cursor;
    const bool have = r.have_blk;
    const unsigned long long blk_idx = r.blk_idx;
    __syncwarp();
    r.cursor = i + 1;
    if (r.mode == RRTFND_STREAM_WORDS) {
        if ((long long)i >= r.n_words) { r.exhausted = true; return 0u; }
        return r.words[i];
    }
    const unsigned long long b = i >> 2;
    if (!have || b != blk_idx) {
        const uint4 o = philox4x32_10(b, r.stream, r.seed);
        r.b0 = o.x; r.b1 = o.y; r.b2 = o.z; r.b3 = o.w;
        r.blk_idx = b;
        r.have_blk = true;
    }
    const unsigned k = (unsigned)i & 3u;
    return k == 0 ? r.b0 : k == 1 ? r.b1 : k == 2 ? r.b2 : r.b3;
}

__device__ __forceinline__ double next_random(WordSrc& r) {
    const uint32_t a = next_word(r) >> 5, b = next_word(r) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);  // all three ops are exact
}

__device__ __forceinline__ uint32_t next_randbelow(WordSrc& r, uint32_t n) {
    const int k = 32 - __clz(n);
    uint32_t v = next_word(r) >> (32 - k);
    while (v >= n && !r.exhausted) v = next_word(r) >> (32 - k);
    return v;
}

}  // namespace rrtfnd
