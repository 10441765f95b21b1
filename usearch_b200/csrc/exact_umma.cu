/*
 *  exact_umma.cu — brute-force search over i8 vectors on Blackwell's 5th-generation tensor cores.
 *
 *  Same contract and the same bits as exact_imma.cu (integer sums are exact in any order: the three i8 metrics are functions
 *  of the integer triple (ab, a2, b2), index_plugins.hpp:1914-1916 / simsimd spatial.h:1880-1972 / dot.h:1749-1775), but the
 *  contraction runs as `tcgen05.mma.cta_group::1.kind::i8` with the accumulators in TENSOR MEMORY and the operands brought
 *  in by TMA tensor copies (SASS: UTCIMMA / LDTM / UTMALDG), instead of warp-level `mma.sync`.
 *
 *  One CTA = one tile of 128 queries against one segment of the stored vectors, walked in tiles of 256 vectors:
 *      warp 0        TMA producer: per k-block of 128 bytes one box of the query tile (128 rows) and one of the vector tile
 *                    (256 rows) into a 4-stage ring of 128B-swizzled shared memory, `full` / `empty` mbarriers
 *      warp 1        MMA issuer: one elected lane issues four M128 x N256 x K32 instructions per k-block into one of TWO
 *                    256-column accumulator buffers in TMEM (512 columns = all of it), `tcgen05.commit` frees the stage and,
 *                    after the last k-block, hands the buffer to the epilogue
 *      warps 2..5    epilogue: thread = TMEM lane = query row. `tcgen05.ld` brings 32 columns at a time into registers; the
 *                    thread turns each integer dot product into the metric's float (i8_distance, shared with the IMMA
 *                    kernel), compares with the row's current worst and — rarely — inserts into that row's k-best list
 *                    under (distance ascending, slot descending). The lists live in SHARED memory (count <= 24; larger
 *                    counts take the mma.sync kernel): with one thread per row an insertion is a chain of dependent loads
 *                    that stalls the whole warp — in global memory (first version: 162 T multiply-adds/s, below the
 *                    mma.sync kernel's 186) that chain cost more than the MMAs. It overlaps with the MMAs of the next tile.
 *  The per-(query, segment) lists are merged by exact_merge_kernel exactly as for the other scan kernels.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "device_index.h"
#include "exact_args.h"
#include "exact_i8.cuh"
#include "warp_primitives.cuh"

namespace usearch_b200 {

namespace {

constexpr int UM_BM = 128;      /* queries per CTA: the M of the instruction, one TMEM lane each */
constexpr int UM_BN = 256;      /* stored vectors per tile: the N of the instruction, one TMEM column each */
constexpr int UM_BK = 128;      /* bytes of K per stage: one 128-byte swizzle row */
constexpr int UM_K = 32;        /* K of one kind::i8 instruction */
constexpr int UM_STAGES = 4;
constexpr int UM_A_BYTES = UM_BM * UM_BK, UM_B_BYTES = UM_BN * UM_BK, UM_STAGE_BYTES = UM_A_BYTES + UM_B_BYTES; /* 48 KB */
constexpr int UM_THREADS = 192; /* warps: 0 TMA, 1 MMA, 2..5 epilogue */
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_KMAX = 24;        /* k-best lists of up to this many entries live in shared memory (row stride 25 words:
                                      conflict-free when every lane touches the same position of its own row) */
constexpr int UM_LIST_STRIDE = UM_KMAX + 1;
constexpr int UM_LIST_BYTES = UM_BM * UM_LIST_STRIDE * 8;

/* instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): dense, no saturate, D = s32 (2) at [4,6),
 * A = B = signed 8 bit (1) at [7,10) / [10,13), both K-major, N >> 3 at [17,23), M >> 4 at [24,29) */
constexpr uint32_t UM_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UM_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

/* shared-memory matrix descriptor of a K-major operand tile written by TMA with the 128-byte swizzle (SmemDescriptor in
 * cute/arch/mma_sm100_desc.hpp): start address >> 4 in [0,14), stride between groups of 8 rows = 1024 B >> 4 in [32,46),
 * version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64); the leading-dimension offset is unused for this layout */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t desc = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    desc |= (uint64_t)(1024u >> 4) << 32;
    desc |= (uint64_t)1 << 46;
    desc |= (uint64_t)2 << 61;
    return desc;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(UM_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, CUtensorMap const* map, uint32_t x, uint32_t y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* one thread, its own list: sorted insert under (distance ascending, slot descending), the order a sequence of
 * sorted_buffer_gt::insert calls in slot order converges to (search_exact_, index.hpp:4251-4268) */
__device__ __noinline__ void list_insert(float* ld, uint32_t* ls, uint32_t& size, uint32_t k, float cd, uint32_t cs) {
    uint32_t pos = size;
    while (pos > 0) { /* entries that sort after the candidate move one place to the right */
        float const d = ld[pos - 1];
        if (d < cd || (d == cd && ls[pos - 1] > cs)) break;
        --pos;
    }
    if (pos >= k) return;
    uint32_t const new_size = size < k ? size + 1 : k;
    for (uint32_t i = new_size - 1; i > pos; --i) { ld[i] = ld[i - 1]; ls[i] = ls[i - 1]; }
    ld[pos] = cd;
    ls[pos] = cs;
    size = new_size;
}

template <uint32_t METRIC, bool SWAP>
__global__ void __launch_bounds__(UM_THREADS, 1) exact_umma_kernel(__grid_constant__ device_index_t const ix,
                                                                   __grid_constant__ exact_args_t const a,
                                                                   __grid_constant__ CUtensorMap const map_queries,
                                                                   __grid_constant__ CUtensorMap const map_vectors) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * UM_STAGES + 4];
    __shared__ uint32_t tmem_base_shared;
    __shared__ int col_b2[2][UM_BN];          /* sum of squares of the tile's vectors, per accumulator buffer */
    __shared__ float col_rn[2][UM_BN];        /* cos: their reciprocal norms */
    __shared__ uint32_t col_mask[2][UM_BN / 32]; /* usable columns: inside the segment and not removed */

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t const stages = (smem_u32(smem_raw) + 1023u) & ~1023u; /* the swizzle atom is 1024 bytes */
    uint32_t const full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[UM_STAGES]);
    uint32_t const tfull0 = smem_u32(&bars[2 * UM_STAGES]), tempty0 = smem_u32(&bars[2 * UM_STAGES + 2]);
    uint32_t const vs = (uint32_t)ix.vec_stride, nkb = (vs + UM_BK - 1) / UM_BK;
    uint32_t const q0 = blockIdx.x * UM_BM;
    uint32_t const seg_lo = blockIdx.y * a.segment_len, seg_hi = min(ix.n, seg_lo + a.segment_len);
    uint32_t const ntiles = seg_hi > seg_lo ? (seg_hi - seg_lo + UM_BN - 1) / UM_BN : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8u * b, 1); mbar_init(tempty0 + 8u * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) { /* this warp owns the tensor memory: all 512 columns, two accumulator buffers */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_shared)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tmem_base = tmem_base_shared;

    if (warp == 0) {
        /* ===== TMA producer ===== */
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = 0; t < ntiles; ++t) {
                uint32_t const tile_base = seg_lo + t * UM_BN;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    uint32_t const sa = stages + stage * UM_STAGE_BYTES, sb = sa + UM_A_BYTES;
                    mbar_expect_tx(full0 + 8u * stage, UM_STAGE_BYTES);
                    tma_load_2d(sa, &map_queries, kb * UM_BK, q0, full0 + 8u * stage);
                    tma_load_2d(sb, &map_vectors, kb * UM_BK, tile_base, full0 + 8u * stage);
                    if (++stage == UM_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        /* ===== MMA issuer ===== */
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = 0; t < ntiles; ++t) {
                uint32_t const buf = t & 1u, buf_phase = (t >> 1) & 1u;
                mbar_wait(tempty0 + 8u * buf, buf_phase ^ 1u); /* the epilogue has drained this buffer */
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t const d_tmem = tmem_base + buf * UM_BN;
                for (uint32_t kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full0 + 8u * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t const sa = stages + stage * UM_STAGE_BYTES, sb = sa + UM_A_BYTES;
                    uint64_t const da = umma_desc(sa), db = umma_desc(sb);
#pragma unroll
                    for (uint32_t k = 0; k < UM_BK / UM_K; ++k) /* 32 bytes further along the swizzled row: +2 in the address field */
                        umma_i8(d_tmem, da + (uint64_t)(k * UM_K >> 4), db + (uint64_t)(k * UM_K >> 4), (kb | k) != 0u);
                    umma_commit(empty0 + 8u * stage); /* the stage is free once these MMAs have read it */
                    if (++stage == UM_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull0 + 8u * buf); /* the accumulators of this tile are complete */
            }
        }
        __syncwarp();
    } else {
        /* ===== epilogue: thread = TMEM lane = query row ===== */
        uint32_t const quarter = (uint32_t)warp & 3u;            /* a warp can only read lanes 32 * (warp % 4) .. +31 */
        uint32_t const row = quarter * 32u + (uint32_t)lane;
        uint32_t const qi = q0 + row;
        bool const live = qi < a.nq;
        int const et = (int)threadIdx.x - 64;                     /* 0..127 among the epilogue threads */
        int const qa2 = (METRIC != METRIC_IP && live) ? a.query_norms[qi] : 0;
        float const qr = METRIC == METRIC_COS ? i8_rnorm(qa2) : 0.f;
        size_t const list = live ? ((size_t)qi * a.segments + blockIdx.y) * a.k : 0;
        float* const ld = reinterpret_cast<float*>(smem_raw + (stages - smem_u32(smem_raw)) + UM_STAGES * UM_STAGE_BYTES) + row * UM_LIST_STRIDE;
        uint32_t* const ls = reinterpret_cast<uint32_t*>(ld - row * UM_LIST_STRIDE + UM_BM * UM_LIST_STRIDE) + row * UM_LIST_STRIDE;
        uint32_t size = 0;
        float worst = 0.f;
        /* filter thresholds while the list is not full: everything passes */
        int thr_i = INT32_MIN;
        float thr_f = -__int_as_float(0x7f800000);
        auto set_thresholds = [&]() { /* the list is full: a column can only enter with d <= worst */
            /* slack: the int -> float conversions and the subtraction round by at most 1.5 ulp of the sum's magnitude
             * (sums beyond 2^24 are not exact in f32); 4 ulps + 4 units are allowed for */
            if constexpr (METRIC == METRIC_IP) { /* d = 1 - float(ab), non-increasing in ab */
                float const t = __fsub_rd(1.0f, worst);
                thr_i = __float2int_rd(t - fabsf(t) * 4.8e-7f) - 4;
            } else if constexpr (METRIC == METRIC_L2SQ) /* d = float(a2 + b2 - 2ab) <= worst  <=>  2ab - b2 >= a2 - floor(worst) (- slack) */
                thr_i = qa2 - (__float2int_ru(worst + fabsf(worst) * 4.8e-7f) + 4);
            else { /* d = 1 - ab*qr*vr <= worst  <=>  ab*vr >= (1 - worst) / qr, lowered by a relative 1e-5 */
                float const base = __fdiv_rn(__fsub_rn(1.0f, worst), qr);
                thr_f = base - fabsf(base) * 1e-5f - 1e-30f;
            }
        };
        for (uint32_t t = 0; t < ntiles; ++t) {
            uint32_t const buf = t & 1u, buf_phase = (t >> 1) & 1u, tile_base = seg_lo + t * UM_BN;
            /* per-column facts of this tile (two columns per thread), while the MMAs run */
            for (int c = et; c < UM_BN; c += 128) {
                uint32_t const slot = tile_base + (uint32_t)c;
                bool usable = slot < seg_hi;
                if (usable && ix.deleted_bits) usable = !((ix.deleted_bits[slot >> 5] >> (slot & 31)) & 1u);
                int const b2 = (METRIC != METRIC_IP && slot < seg_hi) ? a.vector_norms[slot] : 0;
                col_b2[buf][c] = b2;
                col_rn[buf][c] = METRIC == METRIC_COS ? i8_rnorm(b2) : 0.f;
                uint32_t const m = __ballot_sync(0xffffffffu, usable);
                if (lane == 0) col_mask[buf][c >> 5] = m;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory"); /* the four epilogue warps only */
            mbar_wait(tfull0 + 8u * buf, buf_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (uint32_t c0 = 0; c0 < UM_BN; c0 += 32) {
                int v[32];
                tmem_ld32(tmem_base + ((quarter * 32u) << 16) + buf * UM_BN + c0, v);
                /* FILTER, branch-free: which of the 32 columns could still enter this row's list? A conservative test on
                 * the integer dot product (never misses a candidate, may flag a few too many); the first version ran the
                 * exact float test with two branches per element and spent 17 warp instructions per column (ncu). */
                uint32_t pm = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    bool maybe;
                    if constexpr (METRIC == METRIC_IP) maybe = v[j] >= thr_i;
                    else if constexpr (METRIC == METRIC_L2SQ) maybe = 2 * v[j] - col_b2[buf][c0 + j] >= thr_i;
                    else maybe = !(__int2float_rn(v[j]) * col_rn[buf][c0 + j] < thr_f); /* a NaN (zero vector) passes */
                    pm |= maybe ? (1u << j) : 0u;
                }
                pm &= col_mask[buf][c0 >> 5];
                if (live && pm) { /* rare: the exact metric and the sorted insert for the flagged columns */
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (!((pm >> j) & 1u)) continue;
                        float const d = i8_distance<METRIC, SWAP>(v[j], qa2, col_b2[buf][c0 + j], qr, col_rn[buf][c0 + j]);
                        if (size < a.k || !(d > worst)) {
                            list_insert(ld, ls, size, a.k, d, tile_base + c0 + (uint32_t)j);
                            if (size == a.k) {
                                worst = ld[a.k - 1];
                                set_thresholds();
                            }
                        }
                    }
                }
            }
            /* this warp is done reading the buffer: hand it back to the MMA issuer */
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * buf);
            asm volatile("bar.sync 1, 128;" ::: "memory"); /* nobody overwrites col_* of this buffer before all have read it */
        }
        if (live) { /* the row's list -> the per-(query, segment) partial result */
            for (uint32_t i = 0; i < size; ++i) { a.part_d[list + i] = ld[i]; a.part_s[list + i] = ls[i]; }
            a.part_n[(size_t)qi * a.segments + blockIdx.y] = size;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(UM_TMEM_COLS) : "memory");
    }
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point: the library links no libcuda */
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*, cuuint64_t const*,
                                    cuuint32_t const*, cuuint32_t const*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn encode_tiled() {
    static encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<encode_tiled_fn>(p);
    }();
    return fn;
}

/* a row-major byte matrix [rows x row_bytes], rows `pitch` bytes apart, read in boxes of 128 bytes x box_rows with the 128-byte
 * swizzle; bytes and rows outside the matrix arrive as zeros (which add nothing to a dot product) */
bool make_map(CUtensorMap* map, void const* base, uint64_t rows, uint64_t row_bytes, uint64_t pitch, uint32_t box_rows) {
    encode_tiled_fn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t dims[2] = {row_bytes, rows};
    cuuint64_t strides[1] = {pitch};
    cuuint32_t box[2] = {(cuuint32_t)UM_BK, box_rows};
    cuuint32_t elem[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <uint32_t METRIC>
cudaError_t umma_launch_t(device_index_t const& ix, exact_args_t const& a, bool swap, dim3 grid, CUtensorMap const& mq, CUtensorMap const& mv,
                          cudaStream_t stream) {
    size_t const smem = exact_umma_smem_bytes();
    if (swap) {
        cudaError_t e = cudaFuncSetAttribute(exact_umma_kernel<METRIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        exact_umma_kernel<METRIC, true><<<grid, UM_THREADS, smem, stream>>>(ix, a, mq, mv);
    } else {
        cudaError_t e = cudaFuncSetAttribute(exact_umma_kernel<METRIC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        exact_umma_kernel<METRIC, false><<<grid, UM_THREADS, smem, stream>>>(ix, a, mq, mv);
    }
    return cudaGetLastError();
}

} // namespace

size_t exact_umma_smem_bytes() { return (size_t)UM_STAGES * UM_STAGE_BYTES + 1024 + UM_LIST_BYTES; }
int exact_umma_tile_queries() { return UM_BM; }
int exact_umma_tile_vectors() { return UM_BN; }

/* false when the driver cannot encode tensor maps or the operands are not laid out for them: the caller falls back to IMMA */
bool exact_umma_usable(device_index_t const& ix, exact_args_t const& a) {
    return a.k <= (uint32_t)UM_KMAX && encode_tiled() != nullptr && (reinterpret_cast<uintptr_t>(ix.vectors) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.queries) & 15) == 0 &&
           (a.query_stride & 15) == 0 && (ix.vec_stride & 15) == 0;
}

cudaError_t exact_umma_launch(device_index_t const& ix, exact_args_t const& a, bool swap, dim3 grid, cudaStream_t stream) {
    CUtensorMap mq, mv;
    if (!make_map(&mq, a.queries, a.nq, ix.vec_stride, a.query_stride, UM_BM) || !make_map(&mv, ix.vectors, ix.n, ix.vec_stride, ix.vec_stride, UM_BN))
        return cudaErrorInvalidValue;
    switch (ix.metric) {
    case METRIC_IP: return umma_launch_t<METRIC_IP>(ix, a, swap, grid, mq, mv, stream);
    case METRIC_L2SQ: return umma_launch_t<METRIC_L2SQ>(ix, a, swap, grid, mq, mv, stream);
    case METRIC_COS: return umma_launch_t<METRIC_COS>(ix, a, swap, grid, mq, mv, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace usearch_b200
