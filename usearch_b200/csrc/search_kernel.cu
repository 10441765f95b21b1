/*
 *  search_kernel.cu — the batched HNSW search() hot path as ONE persistent, warp-per-query kernel.
 *
 *  What a warp does for one query reproduces, decision for decision, the reference's
 *    index_gt::search            index.hpp:3016-3075
 *    search_for_one_             index.hpp:3963-4003   (greedy descent, levels max..1)
 *    search_to_find_in_base_     index.hpp:4175-4246   (best-first expansion, layer 0)
 *    sorted_buffer_gt::insert    index.hpp:928-939     (`top`, ascending, lower_bound position)
 *    max_heap_gt insert / pop    index.hpp:753-834     (`next`, same sift rules => same tie order)
 *    growing_hash_set_gt::set    index.hpp:1163-1175   (`visits`, exact set semantics)
 *    search_result_t::dump_to    index.hpp:2707-2722   (padding with key 0 / signalling NaN)
 *  so the returned labels are bit-identical to the CPU search at the same ef, ties included.
 *
 *  How the work is laid out on the GPU is new:
 *    - a persistent grid (a multiple of the SM count) pulls query ids from one atomic counter;
 *    - the per-query state lives in shared memory: the query itself (read ~D times), `top`
 *      (sorted, maintained by the whole warp with ballots), the head of the `next` heap (its
 *      deep levels spill to a per-warp slab in HBM), the hop's candidate list;
 *    - `visits` is an open-addressing table in a per-warp slab of HBM driven with atomicCAS at
 *      L2, so that the 32 lanes test-and-set a whole neighbour list at once;
 *    - the M0 neighbour distances of a hop are computed together: LPV lanes per stored vector,
 *      128-bit streaming loads, 8 loads in flight per lane; only the accept/insert replay that
 *      follows is sequential, exactly as the reference's inner loop is.
 */
#include <cuda_runtime.h>

#include "device_index.h"
#include "metrics.cuh"

namespace usearch_b200 {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int THREADS = WARPS_PER_BLOCK * 32;
constexpr int LOADS_IN_FLIGHT = 8;

__device__ __forceinline__ uint32_t hash_slot(uint32_t s) { return s * 0x9E3779B1u; }

/* ---- `next`: binary max-heap on -distance, stored as +distance with reversed compares ------ */

struct heap_t {
    cand_t* smem;
    cand_t* spill;
    uint32_t smem_cap;
    __device__ __forceinline__ cand_t get(uint32_t i) const { return i < smem_cap ? smem[i] : spill[i - smem_cap]; }
    __device__ __forceinline__ void put(uint32_t i, cand_t c) const {
        if (i < smem_cap) smem[i] = c;
        else spill[i - smem_cap] = c;
    }
    /* max_heap_gt::insert_reserved + shift_up (index.hpp:764-770, :808-811). `size` is the size before. */
    __device__ void push(uint32_t size, cand_t c) const {
        uint32_t i = size;
        while (i) {
            uint32_t p = (i - 1) >> 1;
            cand_t pe = get(p);
            if (!(pe.d > c.d)) break; /* less(parent, child) <=> -parent.d < -child.d */
            put(i, pe);
            i = p;
        }
        put(i, c);
    }
    /* max_heap_gt::pop + shift_down (index.hpp:786-794, :819-834). `size` is the size before (>0). */
    __device__ void pop(uint32_t size) const {
        uint32_t n = size - 1;
        if (n == 0) return;
        cand_t last = get(n);
        uint32_t i = 0;
        for (;;) {
            uint32_t l = 2 * i + 1, r = l + 1, best_i = i;
            float best = last.d;
            cand_t le, re;
            if (l < n) { le = get(l); if (best > le.d) { best_i = l; best = le.d; } }
            if (r < n) { re = get(r); if (best > re.d) { best_i = r; } }
            if (best_i == i) break;
            put(i, best_i == l ? le : re);
            i = best_i;
        }
        put(i, last);
    }
};

/* ---- `top`: ascending sorted array, maintained by the whole warp ---------------------------- */

/* sorted_buffer_gt::insert(element, limit) (index.hpp:928-939). Uniform across the warp. */
__device__ __forceinline__ void top_insert(float* td, uint32_t* ts, uint32_t& size, uint32_t limit, float d, uint32_t s,
                                           int lane) {
    uint32_t pos = 0; /* lower_bound: number of stored distances strictly below d */
    for (uint32_t b = 0; b < size; b += 32) {
        uint32_t i = b + lane;
        bool lt = i < size && td[i] < d;
        pos += __popc(__ballot_sync(0xffffffffu, lt));
    }
    if (pos == limit) return;
    bool full = size == limit;
    uint32_t hi = size - (full ? 1u : 0u); /* [pos, hi) shifts one to the right */
    if (hi > pos) {
        for (int b = (int)((hi - 1) & ~31u); b >= (int)(pos & ~31u); b -= 32) {
            uint32_t i = (uint32_t)b + lane;
            bool mv = i >= pos && i < hi;
            float x = 0.f;
            uint32_t y = 0;
            if (mv) { x = td[i]; y = ts[i]; }
            __syncwarp();
            if (mv) { td[i + 1] = x; ts[i + 1] = y; }
            __syncwarp();
        }
    }
    if (lane == 0) { td[pos] = d; ts[pos] = s; }
    size += full ? 0u : 1u;
    __syncwarp();
}

/* ---- distances of a whole candidate list ---------------------------------------------------- */

template <class M>
__device__ __noinline__ void measure_list(device_index_t const& ix, uint4 const* q4, typename M::qconst_t qc,
                                             uint32_t const* cand_s, float* cand_d, uint32_t ncand, int lane) {
    constexpr int LPV = M::LPV, VPP = 32 / LPV;
    int const g = lane / LPV, sub = lane % LPV;
    uint32_t const chunks = ix.chunks16;
    for (uint32_t base = 0; base < ncand; base += VPP) {
        uint32_t c = base + g;
        bool act = c < ncand;
        uint32_t slot = act ? cand_s[c] : 0u;
        uint4 const* v = reinterpret_cast<uint4 const*>(ix.vectors + (size_t)slot * ix.vec_stride);
        typename M::acc_t acc;
        M::init(acc);
        for (uint32_t j0 = sub; j0 < chunks; j0 += LPV * LOADS_IN_FLIGHT) {
            uint4 r[LOADS_IN_FLIGHT];
#pragma unroll
            for (int u = 0; u < LOADS_IN_FLIGHT; ++u) {
                uint32_t j = j0 + u * LPV;
                if (act && j < chunks) r[u] = ldg_stream(v + j);
            }
#pragma unroll
            for (int u = 0; u < LOADS_IN_FLIGHT; ++u) {
                uint32_t j = j0 + u * LPV;
                if (act && j < chunks) M::step(acc, r[u], q4[j]);
            }
        }
        float d = M::finish(acc, qc); /* warp-wide shuffles inside: executed by every lane */
        if (act && sub == 0) cand_d[c] = d;
    }
    __syncwarp();
}

/* ---- one query ------------------------------------------------------------------------------ */

template <class M>
__device__ void search_one(device_index_t const& ix, search_args_t const& a, uint32_t qi, uint8_t* my_smem,
                           uint32_t* visited, cand_t* spill, int lane) {
    uint4* q4 = reinterpret_cast<uint4*>(my_smem);
    float* top_d = reinterpret_cast<float*>(my_smem + a.off_top_d);
    uint32_t* top_s = reinterpret_cast<uint32_t*>(my_smem + a.off_top_s);
    uint32_t* cand_s = reinterpret_cast<uint32_t*>(my_smem + a.off_cand_s);
    float* cand_d = reinterpret_cast<float*>(my_smem + a.off_cand_d);
    heap_t heap{reinterpret_cast<cand_t*>(my_smem + a.off_heap), spill, a.heap_smem_cap};

    uint32_t const k = a.k, ef = a.ef;
    uint32_t top_size = 0, heap_size = 0, computed = 0, cycles = 0, status = STATUS_OK;

    if (ix.n != 0 && k != 0) {
        /* stage the query, zero-padded to whole 16-byte chunks */
        {
            uint8_t const* src = a.queries + (size_t)qi * a.query_stride;
            uint32_t const bpv = ix.bytes_per_vector;
            bool wide = ((reinterpret_cast<size_t>(src) | a.query_stride) & 15) == 0 && a.query_stride >= (uint64_t)ix.chunks16 * 16;
            if (wide) {
                for (uint32_t j = lane; j < ix.chunks16; j += 32) q4[j] = reinterpret_cast<uint4 const*>(src)[j];
            } else {
                uint8_t* dst = reinterpret_cast<uint8_t*>(q4);
                for (uint32_t b = lane; b < ix.chunks16 * 16; b += 32) dst[b] = b < bpv ? src[b] : (uint8_t)0;
            }
        }
        /* visits.clear() */
        {
            uint4 const ones = make_uint4(EMPTY_SLOT, EMPTY_SLOT, EMPTY_SLOT, EMPTY_SLOT);
            uint4* v4 = reinterpret_cast<uint4*>(visited);
            for (uint32_t j = lane; j < a.visited_cap / 4; j += 32) v4[j] = ones;
        }
        __threadfence_block();
        __syncwarp();
        typename M::qconst_t qc = M::prepare(q4, ix.chunks16, lane);
        uint32_t const vmask = a.visited_cap - 1;
        uint32_t visited_count = 0;

        /* ---- search_for_one_: greedy descent (index.hpp:3963-4003) ---- */
        uint32_t closest = ix.entry_slot;
        if (lane == 0) cand_s[0] = closest;
        __syncwarp();
        measure_list<M>(ix, q4, qc, cand_s, cand_d, 1, lane);
        computed += 1;
        float closest_d = cand_d[0];
        __syncwarp();
        for (int level = ix.max_level; level > 0; --level) {
            bool changed;
            do {
                changed = false;
                uint32_t const ubase = ix.upper_base[closest];
                uint32_t const* row = ix.upper + ((size_t)ubase + (uint32_t)(level - 1)) * ix.m_stride;
                uint32_t n = 0;
                for (uint32_t b = 0; b < ix.m; b += 32) {
                    uint32_t i = b + lane;
                    uint32_t s = (i < ix.m && ubase != EMPTY_SLOT) ? row[i] : EMPTY_SLOT;
                    bool valid = s != EMPTY_SLOT;
                    uint32_t bal = __ballot_sync(0xffffffffu, valid);
                    if (valid) cand_s[n + __popc(bal & ((1u << lane) - 1))] = s;
                    n += __popc(bal);
                }
                __syncwarp();
                measure_list<M>(ix, q4, qc, cand_s, cand_d, n, lane);
                computed += n;
                /* sequential `if (d < closest_d)` scan == first occurrence of the strict minimum */
                for (uint32_t b = 0; b < n; b += 32) {
                    uint32_t i = b + lane;
                    float d = i < n ? cand_d[i] : 0.f;
                    bool better = i < n && d < closest_d;
                    /* warp argmin with first-index tie-break */
                    float best = better ? d : __int_as_float(0x7f800000);
                    uint32_t best_i = better ? i : 0xFFFFFFFFu;
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        float od = __shfl_xor_sync(0xffffffffu, best, o);
                        uint32_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                        if (oi != 0xFFFFFFFFu && (best_i == 0xFFFFFFFFu || od < best || (od == best && oi < best_i))) {
                            best = od;
                            best_i = oi;
                        }
                    }
                    if (best_i != 0xFFFFFFFFu) {
                        closest_d = best;
                        closest = cand_s[best_i];
                        changed = true;
                    }
                }
                __syncwarp();
                cycles += 1;
            } while (changed);
        }

        /* ---- search_to_find_in_base_ (index.hpp:4175-4246) ---- */
        if (lane == 0) cand_s[0] = closest;
        __syncwarp();
        measure_list<M>(ix, q4, qc, cand_s, cand_d, 1, lane);
        computed += 1;
        float radius = cand_d[0];
        __syncwarp();
        if (lane == 0) {
            heap.put(0, cand_t{radius, closest});
            atomicCAS(&visited[hash_slot(closest) & vmask], EMPTY_SLOT, closest);
        }
        heap_size = 1;
        visited_count = 1;
        {
            bool allowed = !ix.deleted_bits || !((ix.deleted_bits[closest >> 5] >> (closest & 31)) & 1u);
            if (allowed) {
                if (lane == 0) { top_d[0] = radius; top_s[0] = closest; }
                top_size = 1;
            }
        }
        __syncwarp();

        while (heap_size) {
            cand_t cur = heap.smem[0];
            if (cur.d > radius && top_size == ef) break;
            __syncwarp(); /* every lane holds `cur` before lane 0 rearranges the heap */
            if (lane == 0) heap.pop(heap_size);
            heap_size -= 1;
            cycles += 1;
            __syncwarp();

            /* visits.reserve(): keep the table at most half full so probing terminates quickly */
            if ((visited_count + ix.m0) * 2 > a.visited_cap) { status = STATUS_VISITED_OVERFLOW; break; }

            /* test-and-set every neighbour at once, compact the unseen ones in stored order */
            uint32_t const* row = ix.nbr0 + (size_t)cur.s * ix.m0_stride;
            uint32_t ncand = 0;
            for (uint32_t b = 0; b < ix.m0; b += 32) {
                uint32_t i = b + lane;
                uint32_t s = i < ix.m0 ? __ldg(row + i) : EMPTY_SLOT;
                bool fresh = false;
                if (s != EMPTY_SLOT) {
                    uint32_t h = hash_slot(s) & vmask;
                    for (;;) {
                        uint32_t old = atomicCAS(&visited[h], EMPTY_SLOT, s);
                        if (old == EMPTY_SLOT) { fresh = true; break; }
                        if (old == s) break;
                        h = (h + 1) & vmask;
                    }
                }
                /* a slot listed twice in the same 32-chunk: the first occurrence is the fresh one */
                uint32_t same = __match_any_sync(0xffffffffu, s);
                bool any_fresh = (__ballot_sync(0xffffffffu, fresh) & same) != 0;
                fresh = any_fresh && s != EMPTY_SLOT && (__ffs(same) - 1) == lane;
                uint32_t bal = __ballot_sync(0xffffffffu, fresh);
                if (fresh) cand_s[ncand + __popc(bal & ((1u << lane) - 1))] = s;
                ncand += __popc(bal);
            }
            visited_count += ncand;
            __syncwarp();
            if (ncand == 0) continue;

            measure_list<M>(ix, q4, qc, cand_s, cand_d, ncand, lane);
            computed += ncand;

            /* the reference's sequential accept loop, replayed in stored order */
            for (uint32_t c = 0; c < ncand; ++c) {
                float d = cand_d[c];
                if (top_size < ef || d < radius) {
                    uint32_t s = cand_s[c];
                    if (heap_size >= a.heap_smem_cap + a.heap_spill_cap) { status = STATUS_HEAP_OVERFLOW; break; }
                    if (lane == 0) heap.push(heap_size, cand_t{d, s});
                    heap_size += 1;
                    bool allowed = !ix.deleted_bits || !((ix.deleted_bits[s >> 5] >> (s & 31)) & 1u);
                    if (allowed) {
                        top_insert(top_d, top_s, top_size, ef, d, s, lane);
                        radius = top_d[top_size - 1];
                    }
                    __syncwarp();
                }
            }
            if (status != STATUS_OK) break;
        }
    }

    /* ---- dump_to (index.hpp:2707-2722) ---- */
    __syncwarp();
    uint32_t count = top_size < k ? top_size : k;
    for (uint32_t i = lane; i < k; i += 32) {
        uint64_t key = 0;
        uint32_t bits = SNAN_BITS;
        if (i < count) {
            key = ix.keys[top_s[i]];
            bits = __float_as_uint(top_d[i]);
        }
        a.out_keys[(size_t)qi * k + i] = key;
        reinterpret_cast<uint32_t*>(a.out_dists)[(size_t)qi * k + i] = bits;
    }
    if (lane == 0) {
        a.out_counts[qi] = count;
        if (a.out_computed) a.out_computed[qi] = computed;
        if (a.out_visited) a.out_visited[qi] = cycles;
        a.status[qi] = status;
    }
    __syncwarp();
}

template <class M>
__global__ void __launch_bounds__(THREADS, 4) hnsw_search_kernel(device_index_t ix, search_args_t a) {
    extern __shared__ __align__(16) uint8_t smem[];
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* my_smem = smem + (size_t)warp * a.smem_per_warp;
    uint32_t const gw = blockIdx.x * WARPS_PER_BLOCK + warp;
    uint32_t* visited = a.visited + (size_t)gw * a.visited_cap;
    cand_t* spill = a.heap_spill + (size_t)gw * a.heap_spill_cap;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= a.nq) break;
        uint32_t qi = a.query_list ? a.query_list[item] : item;
        search_one<M>(ix, a, qi, my_smem, visited, spill, lane);
    }
}

/* ---- host-side dispatch --------------------------------------------------------------------- */

template <class M>
static cudaError_t launch_t(device_index_t const& ix, search_args_t const& a, int blocks, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(hnsw_search_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hnsw_search_kernel<M><<<blocks, THREADS, smem, stream>>>(ix, a);
    return cudaGetLastError();
}

template <class M> static cudaError_t occupancy_t(int* blocks_per_sm, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(hnsw_search_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, hnsw_search_kernel<M>, THREADS, smem);
}

#define DISPATCH(FN, ...)                                                                                  \
    switch (ix.scalar) {                                                                                   \
    case SCALAR_F32:                                                                                       \
        if (ix.metric == METRIC_L2SQ) return FN<l2sq_f32_t>(__VA_ARGS__);                                  \
        if (ix.metric == METRIC_IP) return FN<ip_f32_t>(__VA_ARGS__);                                      \
        if (ix.metric == METRIC_COS) return FN<cos_f32_t>(__VA_ARGS__);                                    \
        break;                                                                                             \
    case SCALAR_I8:                                                                                        \
        if (ix.metric == METRIC_L2SQ) return FN<l2sq_i8_t<4>>(__VA_ARGS__);                                \
        if (ix.metric == METRIC_IP) return FN<ip_i8_t<4>>(__VA_ARGS__);                                    \
        if (ix.metric == METRIC_COS) return FN<cos_i8_t<4>>(__VA_ARGS__);                                  \
        break;                                                                                             \
    case SCALAR_B1:                                                                                        \
        if (ix.metric == METRIC_HAMMING) return FN<hamming_b1_t<2>>(__VA_ARGS__);                          \
        if (ix.metric == METRIC_TANIMOTO || ix.metric == METRIC_JACCARD) return FN<tanimoto_b1_t<2>>(__VA_ARGS__); \
        if (ix.metric == METRIC_SORENSEN) return FN<sorensen_b1_t<2>>(__VA_ARGS__);                        \
        break;                                                                                             \
    default: break;                                                                                        \
    }                                                                                                      \
    return cudaErrorInvalidValue;

cudaError_t search_launch(device_index_t const& ix, search_args_t const& a, int blocks, size_t smem, cudaStream_t stream) {
    DISPATCH(launch_t, ix, a, blocks, smem, stream)
}

cudaError_t search_occupancy(device_index_t const& ix, int* blocks_per_sm, size_t smem) {
    DISPATCH(occupancy_t, blocks_per_sm, smem)
}

bool search_supported(uint32_t metric, uint32_t scalar) {
    switch (scalar) {
    case SCALAR_F32:
    case SCALAR_I8: return metric == METRIC_L2SQ || metric == METRIC_IP || metric == METRIC_COS;
    case SCALAR_B1:
        return metric == METRIC_HAMMING || metric == METRIC_TANIMOTO || metric == METRIC_JACCARD || metric == METRIC_SORENSEN;
    default: return false;
    }
}

int search_warps_per_block() { return WARPS_PER_BLOCK; }

} // namespace usearch_b200
