// s2k_count.cuh -- consumer side of the k-min-mer stream: abundance of every distinct k-min-mer hash.
//
// rust-mdbg feeds the iterator's items into a concurrent map keyed by the k-min-mer hash (the reference hints at it:
// "Dashmap level ... kminmer hash collision", src/lib.rs:256-258; KminmerHash compares by hash alone,
// src/kminmer.rs:181-203).  Here the map lives in HBM: an open-addressing table of 64-bit keys with linear probing,
// one atomicCAS to claim a slot, an atomicAdd on its count and an atomicMin on the id of the first item that carried
// the hash (the id leads back to start / end / rev of that occurrence).  Only (hash, count, first id) of the DISTINCT
// hashes leave the device.  Across GPUs the items are partitioned by hash (k_count_hist / k_count_scatter bucket them
// by destination rank for an all-to-all), so that every distinct hash is counted by exactly one rank.
#pragma once
#include "s2k_kernels.cuh"

namespace s2k {

constexpr unsigned long long CT_EMPTY = ~0ull;            // never a stored key: items with this hash go to a side counter
constexpr int CT_MAX_PARTS = 64;

__device__ __forceinline__ uint64_t ct_mix(uint64_t x)      // splitmix64 finalizer: spreads the table and the partition
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
// rank that counts hash h when the items are split over n_parts ranks (host mirror: s2k_count_part)
__device__ __forceinline__ uint32_t ct_part(uint64_t h, uint32_t n_parts)
{
    return __umulhi((uint32_t)(ct_mix(h) >> 32), n_parts);     // floor(top 32 bits * n_parts / 2^32)
}

struct KCArgs {
    const uint64_t *hash;          // n_items
    const uint64_t *id;            // n_items or null: id = id_base + index
    uint64_t n_items, id_base;
    unsigned long long *keys;      // table: capacity slots, CT_EMPTY = free
    uint32_t *cnt;
    unsigned long long *first;     // smallest id seen for the key
    uint64_t mask;                 // capacity - 1 (capacity a power of two)
    unsigned long long *side;      // [0] count of items whose hash is CT_EMPTY, [1] their smallest id, [2] output cursor
};

__global__ void __launch_bounds__(256) k_count_insert(const __grid_constant__ KCArgs A)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n_items; i += stride) {
        const unsigned long long h = A.hash[i];
        const unsigned long long id = A.id ? A.id[i] : A.id_base + i;
        if (h == CT_EMPTY) { atomicAdd(&A.side[0], 1ull); atomicMin(&A.side[1], id); continue; }
        uint64_t slot = ct_mix(h) & A.mask;
        for (;;) {
            unsigned long long cur = A.keys[slot];         // most probes end here: the key is already in the table
            if (cur == CT_EMPTY) cur = atomicCAS(&A.keys[slot], CT_EMPTY, h);
            if (cur == CT_EMPTY || cur == h) {
                atomicAdd(&A.cnt[slot], 1u);
                atomicMin(&A.first[slot], id);
                break;
            }
            slot = (slot + 1) & A.mask;
        }
    }
}

struct KCOut {
    const unsigned long long *keys; const uint32_t *cnt; const unsigned long long *first;
    uint64_t capacity;
    unsigned long long *side;
    uint64_t *out_hash; uint32_t *out_cnt; uint64_t *out_first;
};
// Distinct entries to dense arrays (unordered); one atomic per warp on the output cursor.
__global__ void __launch_bounds__(256) k_count_compact(const __grid_constant__ KCOut A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (A.capacity + stride - 1) / stride;
    for (uint64_t it = 0; it < n_round; ++it) {
        const uint64_t s = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live = s < A.capacity && A.keys[s] != CT_EMPTY;
        const uint32_t m = __ballot_sync(0xffffffffu, live);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&A.side[2], (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) {
            const uint64_t o = base + __popc(m & ((1u << lane) - 1u));
            A.out_hash[o] = A.keys[s]; A.out_cnt[o] = A.cnt[s]; A.out_first[o] = A.first[s];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.side[0]) {    // the one hash that cannot be a key
        const unsigned long long o = atomicAdd(&A.side[2], 1ull);
        A.out_hash[o] = CT_EMPTY; A.out_cnt[o] = (uint32_t)A.side[0]; A.out_first[o] = A.side[1];
    }
}

// Items per destination rank (partition by hash) ...
__global__ void __launch_bounds__(256) k_count_hist(const uint64_t *__restrict__ hash, uint64_t n_items, uint32_t n_parts,
                                                    unsigned long long *__restrict__ part_cnt)
{
    S2K_SHARED uint32_t h[CT_MAX_PARTS];
    if (threadIdx.x < CT_MAX_PARTS) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride)
        atomicAdd(&h[ct_part(hash[i], n_parts)], 1u);
    __syncthreads();
    if (threadIdx.x < n_parts && h[threadIdx.x]) atomicAdd(&part_cnt[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
// ... and the items bucketed by destination: cursor[p] starts at the exclusive prefix of part_cnt.  A CTA claims room for
// its share of every bucket with one atomic per destination, then places its items.
__global__ void __launch_bounds__(256) k_count_scatter(const uint64_t *__restrict__ hash, uint64_t n_items, uint64_t id_base,
                                                       uint32_t n_parts, unsigned long long *__restrict__ cursor,
                                                       uint64_t *__restrict__ out_hash, uint64_t *__restrict__ out_id)
{
    S2K_SHARED uint32_t h[CT_MAX_PARTS];
    S2K_SHARED unsigned long long base[CT_MAX_PARTS];
    const uint64_t per = (n_items + gridDim.x - 1) / gridDim.x;          // a contiguous share per CTA
    const uint64_t i0 = per * blockIdx.x, i1 = min(n_items, i0 + per);
    if (threadIdx.x < CT_MAX_PARTS) h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) atomicAdd(&h[ct_part(hash[i], n_parts)], 1u);
    __syncthreads();
    if (threadIdx.x < n_parts) { base[threadIdx.x] = h[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)h[threadIdx.x]) : 0ull; h[threadIdx.x] = 0; }
    __syncthreads();
    for (uint64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const uint64_t v = hash[i];
        const uint32_t p = ct_part(v, n_parts);
        const uint64_t o = base[p] + atomicAdd(&h[p], 1u);
        out_hash[o] = v; out_id[o] = id_base + i;
    }
}

} // namespace s2k
