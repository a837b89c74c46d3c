// moira_dedup.cu -- dereplication of identical sequences on the device (the reference's --collapse
// dictionary, moira/moira.py:409-411, 459-475): every read gets the index of ONE read with exactly the same sequence
// (its "label": the read that claimed the sequence's slot of an open-addressing table).  Equal label <=> equal sequence,
// byte for byte: a 128-bit hash only selects the candidates, the bytes decide (the table is exact, not probabilistic).
// The reference's choice of representative and order of names (first read with the strictly smallest ee, :466-470) and
// the abundance sort (:492) are integer work on the labels, done by moira_collapse_labels on the host.
//
// Two kernels, one warp per read: seq_hash_kernel (position-keyed 2 x 64-bit sums over 8-byte words, any alignment),
// dedup_insert_kernel (atomicCAS claims a slot; a loser compares hashes, lengths and then the sequences themselves,
// 256 bytes per warp step, and joins or probes on).  Memory-bound; the sequences are read from HBM / L2 twice.
#include <cuda_runtime.h>
#include <stdint.h>

#include "moira_internal.h"

namespace moira {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    return x;
}

struct SeqRef {
    const uint8_t *p;
    uint32_t len;
};

__device__ __forceinline__ SeqRef seq_at(const DedupArgs &a, uint64_t r)
{
    SeqRef s;
    s.p = a.seq + (a.off ? a.off[r] : r * a.stride);
    uint32_t l = a.len ? a.len[r] : a.fixed_len;
    if (a.truncate && l > a.truncate) l = a.truncate;    // the sequence write_results sees: contig[:truncate], moira.py:806-807
    s.len = l;
    return s;
}

// bytes [i, i + 8) of a sequence as a little-endian word, zero beyond `len`; any alignment (two aligned loads + funnel
// shift; the second load may touch up to 7 bytes behind the sequence: callers keep 8 bytes of slack behind their arrays)
__device__ __forceinline__ uint64_t word_at(const uint8_t *p, uint32_t i, uint32_t len)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p) + i;
    const uint32_t mis = (uint32_t)(addr & 7u);
    const uint64_t *q = reinterpret_cast<const uint64_t *>(addr - mis);
    uint64_t w = __ldg(q);
    if (mis) {
        const uint32_t have = 8u - mis;                    // bytes of the word in q[0]
        w >>= 8u * mis;
        if (len - i > have) w |= __ldg(q + 1) << (8u * have);
    }
    const uint32_t rem = len - i;
    if (rem < 8u) w &= (1ull << (8u * rem)) - 1ull;
    return w;
}

__global__ void __launch_bounds__(256) seq_hash_kernel(const DedupArgs a, uint64_t *__restrict__ hash)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < a.n; r += warps) {
        const SeqRef s = seq_at(a, a.base + r);
        uint64_t h1 = 0, h2 = 0;
        for (uint32_t i = lane * 8u; i < s.len; i += 256u) {
            const uint64_t w = word_at(s.p, i, s.len);
            const uint64_t k = (uint64_t)(i >> 3) + 1ull;
            h1 += mix64(w + k * 0x9e3779b97f4a7c15ull);
            h2 += mix64((w ^ (k * 0xc2b2ae3d27d4eb4full)) + 0x165667b19e3779f9ull);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            h1 += __shfl_xor_sync(FULL, h1, o);
            h2 += __shfl_xor_sync(FULL, h2, o);
        }
        if (lane == 0) {
            hash[2 * (a.base + r)] = mix64(h1 ^ s.len);
            hash[2 * (a.base + r) + 1] = mix64(h2 + 0x9e3779b97f4a7c15ull * s.len);
        }
    }
}

__global__ void __launch_bounds__(256) dedup_insert_kernel(const DedupArgs a, const uint64_t *__restrict__ hash, uint32_t *table,
                                                           uint32_t mask, uint32_t *__restrict__ labels)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t rr = warp; rr < a.n; rr += warps) {
        const uint64_t r = a.base + rr;
        const SeqRef s = seq_at(a, r);
        const uint64_t h1 = hash[2 * r], h2 = hash[2 * r + 1];
        uint32_t slot = (uint32_t)h1 & mask;
        for (;;) {
            uint32_t owner = 0;
            if (lane == 0) owner = atomicCAS(&table[slot], EMPTY, (uint32_t)r);
            owner = __shfl_sync(FULL, owner, 0);
            if (owner == EMPTY) { owner = (uint32_t)r; }
            if (owner == (uint32_t)r) {
                if (lane == 0) labels[r] = (uint32_t)r;
                break;
            }
            bool same = hash[2ull * owner] == h1 && hash[2ull * owner + 1] == h2;
            if (same) {
                const SeqRef o = seq_at(a, owner);
                same = o.len == s.len;
                for (uint32_t i0 = 0; same && i0 < s.len; i0 += 256u) {   // the bytes decide
                    const uint32_t i = i0 + lane * 8u;
                    const bool eq = i >= s.len || word_at(s.p, i, s.len) == word_at(o.p, i, s.len);
                    same = __all_sync(FULL, eq);
                }
            }
            if (same) {
                if (lane == 0) labels[r] = owner;
                break;
            }
            slot = (slot + 1u) & mask;
        }
    }
}

}  // namespace

int launch_dedup(const DedupArgs &a, uint64_t *d_hash, uint32_t *d_table, uint32_t table_mask, uint32_t *d_labels, const LaunchCfg &cfg)
{
    if (!a.n) return 0;
    const int grid = cfg.sm_count * 8;
    seq_hash_kernel<<<grid, 256, 0, cfg.stream>>>(a, d_hash);
    dedup_insert_kernel<<<grid, 256, 0, cfg.stream>>>(a, d_hash, d_table, table_mask, d_labels);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace moira
