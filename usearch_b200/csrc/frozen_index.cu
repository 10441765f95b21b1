/*
 *  frozen_index.cu — host side of the GPU search backend: parse the reference's v2 serialisation,
 *  lay the graph out as flat arrays in HBM, plan launches, run batches, retry scratch overflows.
 *
 *  Reference behaviour mirrored here (file:line under /root/reference/include/usearch):
 *    index_dense.hpp:1084-1188  load_from_stream: [u32 rows, u32 cols][matrix][64-byte head][graph]
 *    index_dense.hpp:42-79      index_dense_head_t field order
 *    index.hpp:3322-3382        graph: 40-byte header | int16 levels | node tapes
 *    index.hpp:2116-2195        node tape: key u64 | level i16 | {u32 n, slot[M0]} | level x {u32 n, slot[M]}
 *    index.hpp:3016-3075        expansion = max(config.expansion or 64, wanted)
 *    index_plugins.hpp:1105-1224 query casts
 */
#include "frozen_index.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace usearch_b200 {

namespace {

uint64_t rd_u64(uint8_t const* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
uint32_t rd_u32(uint8_t const* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
int16_t rd_i16(uint8_t const* p) { int16_t v; std::memcpy(&v, p, 2); return v; }
uint32_t ceil2(uint64_t v) { uint64_t r = 1; while (r < v) r <<= 1; return (uint32_t)std::min<uint64_t>(r, 1ull << 31); }
uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

char const* cuda_error(cudaError_t e) {
    if (e == cudaSuccess) return nullptr;
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return "Out of GPU memory!";
    static thread_local char message[160];
    std::snprintf(message, sizeof(message), "CUDA failure: %s", cudaGetErrorString(e));
    return message;
}

#define CU(call)                                             \
    do {                                                     \
        if (char const* err_ = cuda_error((call))) return err_; \
    } while (0)

} // namespace

int default_device() {
    if (char const* dev = std::getenv("USEARCH_B200_DEVICE")) return std::atoi(dev);
    if (char const* rank = std::getenv("LOCAL_RANK")) return std::atoi(rank);
    return 0;
}

size_t bits_per_scalar(uint32_t s) { /* index_plugins.hpp:237-257 */
    switch (s) {
    case SCALAR_B1: return 1;
    case SCALAR_I8: return 8;
    case SCALAR_F16: case SCALAR_BF16: return 16;
    case SCALAR_F32: return 32;
    case SCALAR_F64: return 64;
    default: return 0;
    }
}

frozen_index_t::~frozen_index_t() {
    for (pending_search_t& p : pending) pending_free.push_back(p.status);
    for (device_buffer_t<uint32_t>* b : pending_free) { b->release(); delete b; }
    leave_shards();
    release_device();
    if (stream) cudaStreamDestroy(stream);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    phase_cycles.release();
    build.release();
    cast_stage.release();
    exact_scratch.release();
    visit_log.release();
    visited.release(); work_counter.release(); status.release(); counts.release(); computed.release();
    cycles.release(); retry_list.release(); heap_spill.release(); queries.release(); out_keys.release();
    allowed_keys.release(); allow_bits.release();
    out_dists.release(); h_queries.release(); h_keys.release(); h_dists.release(); h_counts.release();
    h_computed.release(); h_cycles.release(); h_status.release();
}

void frozen_index_t::release_device() {
    for (void*& p : dev_allocs) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    d = device_index_t{};
    hbm_bytes = 0;
    loaded = false;
    size = 0;
    count_deleted = 0;
    capacity = 0;
    upper_capacity = 0;
    upper_rows = 0;
    visited_zeroed_words = 0;
    levels.clear();
    host_keys.clear();
    key_map.clear();
}

char const* frozen_index_t::ensure_context() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return "No CUDA device: the B200 search backend has no CPU fallback";
    }
    CU(cudaSetDevice(device));
    if (!stream) {
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&ev_begin));
        CU(cudaEventCreate(&ev_end));
        CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    }
    return nullptr;
}

char const* frozen_index_t::counts_reserve_all(size_t nq) {
    if (char const* e = counts.reserve(nq)) return e;
    if (char const* e = computed.reserve(nq)) return e;
    if (char const* e = cycles.reserve(nq)) return e;
    if (char const* e = h_counts.reserve(nq)) return e;
    if (char const* e = h_computed.reserve(nq)) return e;
    if (char const* e = h_cycles.reserve(nq)) return e;
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  v2 blob -> HBM                                                                                */
/* ---------------------------------------------------------------------------------------------- */

char const* frozen_index_t::load_blob(uint8_t const* blob, size_t length) {
    if (char const* e = ensure_context()) return e;
    release_device();

    uint8_t const* p = blob;
    uint8_t const* const end = blob + length;
    if (length < 8 + 64 + 40) return "File is corrupted and lacks matrix dimensions";
    uint64_t const rows = rd_u32(p), cols = rd_u32(p + 4);
    p += 8;
    if ((uint64_t)(end - p) < rows * cols + 64 + 40) return "File is corrupted and lacks a header";
    uint8_t const* const matrix = p;
    p += rows * cols;
    if (std::memcmp(p, "usearch", 7) != 0) return "Magic header mismatch - the file isn't an index";
    uint16_t version_major;
    std::memcpy(&version_major, p + 7, 2);
    if (version_major != 2) return "File format may be different, please rebuild";
    uint32_t const head_metric = p[13], head_scalar = p[14], head_key = p[15], head_slot = p[16];
    if (head_key != 14 /* u64_k */) return "Key type doesn't match, consider rebuilding";
    if (head_slot != 15 /* u32_k */) return "Slot type doesn't match, consider rebuilding";
    uint64_t const count_present = rd_u64(p + 17), deleted = rd_u64(p + 25), dims = rd_u64(p + 33);
    bool const head_multi = p[41] != 0;
    p += 64;
    (void)count_present;
    if (!search_supported(head_metric, head_scalar))
        return "This metric / scalar kind has no sm_100a kernel and the backend has no CPU fallback";
    size_t const bpv = (dims * bits_per_scalar(head_scalar) + 7) / 8;
    if (rows && cols != bpv) return "Matrix columns do not match bytes per vector";

    uint64_t const n = rd_u64(p), m = rd_u64(p + 8), m0 = rd_u64(p + 16), max_level = rd_u64(p + 24), entry = rd_u64(p + 32);
    p += 40;
    if (n != rows) return "Index size and the number of vectors doesn't match";
    if (n >= 0xFFFFFFFFull) return "Too many entries for 32-bit slots";
    if (n && (m < 2 || m0 < 2)) return "Connectivity is too low";
    if ((uint64_t)(end - p) < n * 2) return "File is corrupted and can't fit all the levels";
    uint8_t const* const levels_bytes = p;
    p += n * 2;

    /* host fields are committed only when the whole file has been accepted (see `commit` below) */
    auto commit = [&](device_index_t const& accepted, size_t rows_in_upper) {
        metric = head_metric;
        scalar = head_scalar;
        dimensions = dims;
        connectivity = m;
        connectivity_base = m0;
        multi = head_multi;
        size = n;
        count_deleted = deleted;
        capacity = n;
        upper_rows = rows_in_upper;
        upper_capacity = std::max<size_t>(rows_in_upper, 1);
        d = accepted;
        loaded = true;
    };
    if (n && max_level > 0x7FFF) return "File is corrupted: level out of range";

    device_index_t ix;
    ix.n = (uint32_t)n;
    ix.m = (uint32_t)m;
    ix.m0 = (uint32_t)m0;
    ix.m_stride = round_up((uint32_t)m, 4);
    ix.m0_stride = round_up((uint32_t)m0, 4);
    ix.entry_slot = (uint32_t)entry;
    ix.max_level = (int32_t)max_level;
    ix.dims = (uint32_t)dims;
    ix.bytes_per_vector = (uint32_t)bpv;
    ix.vec_stride = round_up((uint32_t)bpv, 16);
    ix.chunks16 = (uint32_t)(ix.vec_stride / 16);
    ix.metric = head_metric;
    ix.scalar = head_scalar;
    if (n == 0) {
        commit(ix, 0);
        upper_capacity = 0;
        return nullptr;
    }
    if (entry >= n) return "File is corrupted: entry slot out of range";

    /* pass 1: levels, upper row offsets, tape offsets */
    levels.resize(n);
    host_keys.resize(n);
    struct drop_host_state_t { /* a rejected file leaves no trace in the handle */
        frozen_index_t* self;
        bool armed = true;
        ~drop_host_state_t() { if (armed) { self->levels.clear(); self->host_keys.clear(); self->release_device(); } }
    } drop_host_state{this};
    std::vector<uint32_t> upper_base(n);
    uint64_t upper_rows = 0;
    size_t const nb = m * 4 + 4, nb0 = m0 * 4 + 4;
    {
        uint8_t const* q = p;
        for (uint64_t i = 0; i < n; ++i) {
            int16_t level = rd_i16(levels_bytes + 2 * i);
            if (level < 0) return "File is corrupted: negative level";
            levels[i] = level;
            upper_base[i] = level ? (uint32_t)upper_rows : EMPTY_SLOT;
            upper_rows += (uint64_t)level;
            size_t node_bytes = 10 + nb0 + nb * (size_t)level;
            if ((size_t)(end - q) < node_bytes) return "File is corrupted and can't fit all the nodes";
            q += node_bytes;
        }
        if (upper_rows >= 0xFFFFFFFFull) return "Too many upper-level rows";
        /* the descent starts on `max_level` at the entry point (index.hpp:3963-3975): it must own that many rows */
        if ((uint64_t)levels[entry] != max_level) return "File is corrupted: entry point and top level disagree";
    }

    /* device allocations */
    uint8_t* d_vectors = nullptr;
    uint64_t* d_keys = nullptr;
    uint32_t *d_nbr0 = nullptr, *d_upper_base = nullptr, *d_upper = nullptr, *d_deleted = nullptr;
    size_t const bytes_vectors = (size_t)n * ix.vec_stride, bytes_keys = (size_t)n * 8,
                 bytes_nbr0 = (size_t)n * ix.m0_stride * 4, bytes_ub = (size_t)n * 4,
                 bytes_upper = std::max<size_t>(upper_rows, 1) * ix.m_stride * 4, bytes_deleted = ((size_t)n + 31) / 32 * 4;
    CU(cudaMalloc(&d_vectors, bytes_vectors)); dev_allocs[0] = d_vectors;
    CU(cudaMalloc(&d_keys, bytes_keys)); dev_allocs[1] = d_keys;
    CU(cudaMalloc(&d_nbr0, bytes_nbr0)); dev_allocs[2] = d_nbr0;
    CU(cudaMalloc(&d_upper_base, bytes_ub)); dev_allocs[3] = d_upper_base;
    CU(cudaMalloc(&d_upper, bytes_upper)); dev_allocs[4] = d_upper;
    hbm_bytes = bytes_vectors + bytes_keys + bytes_nbr0 + bytes_ub + bytes_upper;

    /* vectors: slot-major matrix, rows padded to 16 bytes */
    if (ix.vec_stride == bpv) {
        CU(cudaMemcpy(d_vectors, matrix, bytes_vectors, cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemset(d_vectors, 0, bytes_vectors));
        CU(cudaMemcpy2D(d_vectors, ix.vec_stride, matrix, bpv, bpv, n, cudaMemcpyHostToDevice));
    }

    /* pass 2: node tapes -> keys, nbr0 rows, upper rows; converted and uploaded in chunks */
    size_t const chunk_nodes = 1u << 18;
    std::vector<uint64_t> h_keys(std::min<size_t>(n, chunk_nodes));
    std::vector<uint32_t> h_nbr0(std::min<size_t>(n, chunk_nodes) * ix.m0_stride);
    std::vector<uint32_t> h_upper;
    std::vector<uint32_t> h_deleted(((size_t)n + 31) / 32, 0u);
    bool any_deleted = false;
    uint8_t const* q = p;
    for (uint64_t begin = 0; begin < n; begin += chunk_nodes) {
        uint64_t const stop = std::min<uint64_t>(n, begin + chunk_nodes);
        uint64_t const first_upper_row = [&] { for (uint64_t i = begin; i < stop; ++i) if (levels[i]) return (uint64_t)upper_base[i]; return upper_rows; }();
        uint64_t chunk_upper_rows = 0;
        for (uint64_t i = begin; i < stop; ++i) chunk_upper_rows += (uint64_t)levels[i];
        h_upper.assign(chunk_upper_rows * ix.m_stride, EMPTY_SLOT);
        std::fill(h_nbr0.begin(), h_nbr0.begin() + (stop - begin) * ix.m0_stride, EMPTY_SLOT);
        uint64_t row = 0;
        for (uint64_t i = begin; i < stop; ++i) {
            uint64_t key = rd_u64(q);
            h_keys[i - begin] = key;
            host_keys[i] = key;
            if (key == free_key) { h_deleted[i >> 5] |= 1u << (i & 31); any_deleted = true; }
            uint8_t const* list = q + 10;
            uint32_t c0 = std::min<uint32_t>(rd_u32(list), (uint32_t)m0);
            uint32_t* dst0 = h_nbr0.data() + (i - begin) * ix.m0_stride;
            /* A self-link or a repeated slot in a layer-0 list can never pass `visits.set()` in
             * search_to_find_in_base_ (index.hpp:4221): the expanded node and the first occurrence are
             * already marked. Dropping them here changes nothing observable and lets the kernel issue
             * every visited test of a row at once. */
            uint32_t kept = 0;
            for (uint32_t j = 0; j < c0; ++j) {
                uint32_t s = rd_u32(list + 4 + 4 * j);
                if (s >= n) return "File is corrupted: neighbour slot out of range";
                bool drop = s == (uint32_t)i;
                for (uint32_t t = 0; t < kept && !drop; ++t) drop = dst0[t] == s;
                if (!drop) dst0[kept++] = s;
            }
            list += nb0;
            for (int16_t l = 0; l < levels[i]; ++l, ++row, list += nb) {
                uint32_t c = std::min<uint32_t>(rd_u32(list), (uint32_t)m);
                uint32_t* dst = h_upper.data() + row * ix.m_stride;
                for (uint32_t j = 0; j < c; ++j) {
                    uint32_t s = rd_u32(list + 4 + 4 * j);
                    if (s >= n) return "File is corrupted: neighbour slot out of range";
                    /* a search on level l+1 reads row l+1 of every member it reaches: the member must have it */
                    if (levels[s] < l + 1) return "File is corrupted: link to a member that is absent from that level";
                    dst[j] = s;
                }
            }
            q = list;
        }
        CU(cudaMemcpy(d_keys + begin, h_keys.data(), (stop - begin) * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_nbr0 + begin * ix.m0_stride, h_nbr0.data(), (stop - begin) * ix.m0_stride * 4, cudaMemcpyHostToDevice));
        if (chunk_upper_rows)
            CU(cudaMemcpy(d_upper + first_upper_row * ix.m_stride, h_upper.data(), chunk_upper_rows * ix.m_stride * 4,
                          cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(d_upper_base, upper_base.data(), bytes_ub, cudaMemcpyHostToDevice));
    if (any_deleted) {
        CU(cudaMalloc(&d_deleted, bytes_deleted)); dev_allocs[5] = d_deleted;
        CU(cudaMemcpy(d_deleted, h_deleted.data(), bytes_deleted, cudaMemcpyHostToDevice));
        hbm_bytes += bytes_deleted;
    }
    ix.vectors = d_vectors;
    ix.keys = d_keys;
    ix.nbr0 = d_nbr0;
    ix.upper_base = d_upper_base;
    ix.upper = d_upper;
    ix.deleted_bits = d_deleted;
    if (search_needs_norms(head_metric, head_scalar)) {
        float* d_norms = nullptr;
        CU(cudaMalloc(&d_norms, (size_t)n * 4)); dev_allocs[6] = d_norms;
        hbm_bytes += (size_t)n * 4;
        CU(search_compute_norms(ix, d_norms, stream));
        CU(cudaStreamSynchronize(stream));
        ix.norms = d_norms;
    }
    drop_host_state.armed = false;
    commit(ix, upper_rows);
    return nullptr;
}

size_t frozen_index_t::serialized_length() const {
    size_t const nb = connectivity * 4 + 4, nb0 = connectivity_base * 4 + 4;
    size_t total = 8 + size * d.bytes_per_vector + 64 + 40 + size * 2;
    for (size_t i = 0; i < size; ++i) total += 10 + nb0 + nb * (size_t)levels[i];
    return total;
}

/* Re-serialise the frozen index into the reference's v2 format (index_dense.hpp:994-1062,
 * index.hpp:3276-3317) by downloading the SoA arrays. */
char const* frozen_index_t::save_blob(uint8_t* out, size_t length) const {
    if (length < serialized_length()) return "Failed to serialize into stream";
    CU(cudaSetDevice(device));
    size_t const n = size, bpv = d.bytes_per_vector;
    uint8_t* p = out;
    uint32_t dims32[2] = {(uint32_t)n, (uint32_t)bpv};
    std::memcpy(p, dims32, 8);
    p += 8;
    if (n) {
        if (d.vec_stride == bpv) CU(cudaMemcpy(p, d.vectors, n * bpv, cudaMemcpyDeviceToHost));
        else CU(cudaMemcpy2D(p, bpv, d.vectors, d.vec_stride, bpv, n, cudaMemcpyDeviceToHost));
    }
    p += n * bpv;
    std::memset(p, 0, 64);
    std::memcpy(p, "usearch", 7);
    uint16_t version[3] = {2, 21, 0};
    std::memcpy(p + 7, version, 6);
    p[13] = (uint8_t)metric;
    p[14] = (uint8_t)scalar;
    p[15] = 14; /* u64 keys */
    p[16] = 15; /* u32 slots */
    uint64_t present = n - count_deleted, deleted = count_deleted, dims = dimensions;
    std::memcpy(p + 17, &present, 8);
    std::memcpy(p + 25, &deleted, 8);
    std::memcpy(p + 33, &dims, 8);
    p[41] = multi ? 1 : 0;
    p += 64;
    uint64_t header[5] = {n, connectivity, connectivity_base, (uint64_t)d.max_level, d.entry_slot};
    std::memcpy(p, header, 40);
    p += 40;
    if (!n) return nullptr;
    std::memcpy(p, levels.data(), n * 2);
    p += n * 2;
    std::vector<uint64_t> h_keys(n);
    std::vector<uint32_t> h_nbr0((size_t)n * d.m0_stride), h_ub(n);
    size_t upper_rows = 0;
    for (size_t i = 0; i < n; ++i) upper_rows += (size_t)levels[i];
    std::vector<uint32_t> h_upper(std::max<size_t>(upper_rows, 1) * d.m_stride);
    CU(cudaMemcpy(h_keys.data(), d.keys, n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_nbr0.data(), d.nbr0, h_nbr0.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_ub.data(), d.upper_base, n * 4, cudaMemcpyDeviceToHost));
    if (upper_rows) CU(cudaMemcpy(h_upper.data(), d.upper, upper_rows * d.m_stride * 4, cudaMemcpyDeviceToHost));
    size_t const nb = connectivity * 4 + 4, nb0 = connectivity_base * 4 + 4;
    auto write_list = [&](uint32_t const* row, uint32_t cap, size_t bytes) {
        std::memset(p, 0, bytes);
        uint32_t c = 0;
        while (c < cap && row[c] != EMPTY_SLOT) ++c;
        std::memcpy(p, &c, 4);
        std::memcpy(p + 4, row, (size_t)c * 4);
        p += bytes;
    };
    for (size_t i = 0; i < n; ++i) {
        std::memcpy(p, &h_keys[i], 8);
        std::memcpy(p + 8, &levels[i], 2);
        p += 10;
        write_list(h_nbr0.data() + i * d.m0_stride, d.m0, nb0);
        for (int16_t l = 0; l < levels[i]; ++l) write_list(h_upper.data() + ((size_t)h_ub[i] + l) * d.m_stride, d.m, nb);
    }
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  launch planning                                                                               */
/* ---------------------------------------------------------------------------------------------- */

char const* frozen_index_t::plan(uint32_t k, uint32_t visited_cap_override, launch_plan_t& pl, uint32_t ef_override) const {
    uint32_t ef = (uint32_t)(expansion_search ? expansion_search : 64); /* index.hpp:3029-3030 */
    ef = std::max(ef, k);                                               /* index.hpp:3052 */
    if (ef_override) ef = ef_override; /* the builder searches with expansion_add (index.hpp:2854) */
    pl.ef = ef;
    /* `visits` covers every slot the arrays have room for, so that its size does not change while members are added */
    uint64_t const visit_slots = std::max<uint64_t>(capacity, d.n);
    uint32_t const list_cap = round_up(std::max(d.m0, d.m), 32);
    /* per-warp (= per-CTA) shared memory: query | top | candidates | mbarriers | TMA slots | heap head */
    uint32_t off = 0;
    off += d.chunks16 * 16;          /* query */
    uint32_t const top_smem = ef > 256 ? round_up(ef * 4, 16) : 0; /* ef <= 256: `top` lives in registers */
    pl.off_top_d = off; off += top_smem;
    pl.off_top_s = off; off += top_smem;
    pl.off_cand_s = off; off += list_cap * 4;
    pl.off_cand_d = off; off += list_cap * 4;
    pl.off_bars = off; off += 256; /* 32 mbarriers */
    off = round_up(off, 128);
    pl.off_stage = off;
    int const slots = search_stage_slots(d); /* slots of one set: 32 / LPV */
    /* an SM has 228 KB of shared memory and charges 1 KB per resident CTA on top of its request */
    size_t const smem_sm = 228 * 1024, cta_tax = 1024, smem_cta_max = 227 * 1024;
    uint32_t const min_heap = 128 * 8;
    int const forced_sets = tune.stage_sets;
    pl.stage_sets = 1;
    pl.stage_stride = 0;
    if (slots) {
        /* slot stride = 16*LPV mod 128 bytes: the lanes of a quarter-warp then read disjoint banks */
        pl.stage_stride = round_up(d.chunks16 * 16, 128) + search_stage_pad(d);
        /* double-buffer the slots when at least 4 warps per SM still fit */
        size_t const two = off + 2 * (size_t)slots * pl.stage_stride + min_heap + cta_tax;
        pl.stage_sets = smem_sm / two >= 4 ? 2u : 1u;
        /* short vectors of the 16-warp kernels: resident warps beat double buffering (search_kernel.cu, dispatch) */
        if (search_single_stage_set(d)) pl.stage_sets = 1;
        if (forced_sets == 1 || forced_sets == 2) pl.stage_sets = (uint32_t)forced_sets;
    }
    off += (uint32_t)slots * pl.stage_sets * pl.stage_stride;
    pl.off_heap = off;
    uint32_t const fixed = off;
    if (fixed + min_heap > smem_cta_max) return "Expansion or dimensionality too large for on-chip state";
    int const forced_warps = tune.warps_per_sm;
    uint32_t warps_sm = (uint32_t)std::min<size_t>(smem_sm / (fixed + min_heap + cta_tax), (size_t)search_max_warps_per_sm(d));
    if (forced_warps > 0) warps_sm = std::min<uint32_t>(warps_sm, (uint32_t)forced_warps);
    warps_sm = std::max(warps_sm, 1u);
    uint32_t budget = (uint32_t)(smem_sm / warps_sm - cta_tax);
    budget = std::min<uint32_t>(budget, (uint32_t)smem_cta_max);
    uint32_t heap_bytes = std::min<uint32_t>((budget - fixed) & ~15u, 4096 * 8);
    pl.heap_smem_cap = heap_bytes / 8; /* even: heap_bytes is a multiple of 16 */
    pl.smem_per_warp = fixed + pl.heap_smem_cap * 8;
    pl.smem_per_block = pl.smem_per_warp;
    pl.warps_per_sm_target = warps_sm;

    /* scratch per warp. `scale` (1, 8, 64, ...) grows it for the retry of overflowed queries. */
    uint64_t const scale = visited_cap_override ? visited_cap_override : 1;
    uint64_t const enough = (uint64_t)2 * (visit_slots + d.m0 + 1); /* a hash table this large can never overflow */
    static int const forced = [] { /* test hook: USEARCH_B200_VISITED=hash|bitmap|bitmap_log */
        char const* v = std::getenv("USEARCH_B200_VISITED");
        return !v ? 0 : (std::strcmp(v, "hash") == 0 ? 1 : (std::strcmp(v, "bitmap") == 0 ? 2 : (std::strcmp(v, "bitmap_log") == 0 ? 3 : 0)));
    }();
    static uint64_t const shrink = [] { /* test hook: start with undersized scratch to exercise the retry path */
        char const* v = std::getenv("USEARCH_B200_SCRATCH_SHRINK");
        return v && std::atoi(v) > 0 ? (uint64_t)std::atoi(v) : (uint64_t)1;
    }();
    uint64_t const bitmap_words = round_up((uint32_t)((visit_slots + 31) / 32), 4);
    uint64_t const max_warps_guess = (uint64_t)sm_count * 32;
    bool const bitmaps_fit = bitmap_words * 4 * std::min<uint64_t>(max_warps_guess, (uint64_t)sm_count * pl.warps_per_sm_target) <= BITMAP_SCRATCH_BUDGET;
    if (forced == 2 || forced == 3 || (forced == 0 && bitmaps_fit)) {
        /* BITMAP visits: one bit per slot, exact, never overflows */
        pl.visited_bitmap_words = (uint32_t)bitmap_words;
        pl.visited_cap = 0;
        pl.visit_log_cap = (forced == 3 || (forced == 0 && visit_slots > BITMAP_WIPE_MAX_SLOTS))
                               ? (uint32_t)std::max<uint64_t>(32768 / shrink, 64) : 0u;
        uint64_t spill = std::max<uint64_t>(1024, (uint64_t)8 * ef) * scale / shrink;
        spill = std::max<uint64_t>(spill, 16);
        pl.heap_spill_cap = (uint32_t)std::min<uint64_t>(spill, visit_slots + 1);
        pl.maxed = pl.heap_spill_cap >= visit_slots;
    } else {
        pl.visited_bitmap_words = 0;
        pl.visit_log_cap = 0;
        uint64_t want = std::max<uint64_t>((uint64_t)2 * ef * d.m0 * scale, 2048) / shrink;
        pl.visited_cap = std::max<uint32_t>(ceil2(std::min<uint64_t>(want, enough)), 64);
        /* pushes <= visited entries <= cap/2, so this spill can not overflow before `visits` does */
        pl.heap_spill_cap = pl.visited_cap / 2;
        pl.maxed = pl.visited_cap >= enough;
    }

    int per_sm = 0;
    CU(search_occupancy(d, &per_sm, pl.smem_per_block));
    if (per_sm < 1) return "Kernel does not fit on an SM";
    per_sm = std::min<int>(per_sm, (int)pl.warps_per_sm_target);
    pl.blocks = per_sm * sm_count;
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  batched search on device buffers                                                              */
/* ---------------------------------------------------------------------------------------------- */

/* scratch for `warps` resident warps under plan `pl`, and the launch arguments that describe it */
char const* frozen_index_t::prepare_launch(launch_plan_t const& pl, size_t warps, search_args_t& a, cudaStream_t s) {
    if (char const* e = work_counter.reserve(2)) return e;
    size_t const words = warps * pl.visited_words_per_warp();
    bool const grown = words > visited.capacity;
    if (char const* e = visited.reserve(words)) return e;
    if (grown) visited_zeroed_words = 0;
    if (pl.visit_log_cap) { /* logged bitmaps rely on an all-zero slab between queries */
        if (char const* e = visit_log.reserve(warps * pl.visit_log_cap)) return e;
        if (visited_zeroed_words < words) {
            if (cudaMemsetAsync(visited.ptr, 0, words * 4, s) != cudaSuccess) return "CUDA failure: memset";
            visited_zeroed_words = words;
        }
    } else
        visited_zeroed_words = 0; /* wiped per query with other contents in between */
    if (char const* e = heap_spill.reserve(warps * pl.heap_spill_cap)) return e;
    a = search_args_t{};
    a.ef = pl.ef;
    a.work_counter = work_counter.ptr;
    a.visited = visited.ptr;
    a.visited_cap = pl.visited_cap;
    a.visited_bitmap_words = pl.visited_bitmap_words;
    a.visit_log = pl.visit_log_cap ? visit_log.ptr : nullptr;
    a.visit_log_cap = pl.visit_log_cap;
    a.heap_spill = heap_spill.ptr;
    a.heap_spill_cap = pl.heap_spill_cap;
    a.heap_smem_cap = pl.heap_smem_cap;
    a.smem_per_warp = pl.smem_per_warp;
    a.off_top_d = pl.off_top_d; a.off_top_s = pl.off_top_s; a.off_cand_s = pl.off_cand_s;
    a.off_cand_d = pl.off_cand_d; a.off_heap = pl.off_heap;
    a.off_bars = pl.off_bars; a.off_stage = pl.off_stage; a.stage_stride = pl.stage_stride;
    a.stage_sets = pl.stage_sets;
    return nullptr;
}

char const* frozen_index_t::search_device(void const* d_queries, size_t nq, size_t stride, size_t k, uint64_t* d_keys,
                                          float* d_dists, uint32_t* d_counts, uint32_t* d_computed, uint32_t* d_cycles,
                                          cudaStream_t s, bool defer) {
    if (nq == 0 || k == 0) return nullptr;
    if (nq > 0x7FFFFFFFull) return "Too many queries in one batch";
    if (char const* e = ensure_context()) return e;
    if (!loaded || d.n == 0) { /* no matches, no error (index.hpp:3036-3037) */
        CU(search_fill_empty(d_keys, d_dists, d_counts, d_computed, d_cycles, nq, k, s));
        return nullptr;
    }
    launch_plan_t pl;
    if (char const* e = plan((uint32_t)k, 0, pl)) return e;
    /* a small batch does not need the whole grid */
    int const wpb = search_warps_per_block();
    int blocks = (int)std::min<size_t>((size_t)pl.blocks, (nq + wpb - 1) / wpb);
    size_t warps = (size_t)blocks * wpb;

    if (char const* e = h_status.reserve(nq)) return e;
    uint32_t* status_ptr = nullptr;
    if (defer) { /* every batch in flight owns its status words until search_finish has looked at them */
        if (pending_free.empty()) pending_free.emplace_back(new device_buffer_t<uint32_t>());
        device_buffer_t<uint32_t>* buf = pending_free.back();
        pending_free.pop_back();
        if (char const* e = buf->reserve(nq)) { pending_free.push_back(buf); return e; }
        pending.push_back(pending_search_t{buf, search_args_t{}, pl.maxed, s});
        status_ptr = buf->ptr;
    } else {
        if (char const* e = status.reserve(nq)) return e;
        status_ptr = status.ptr;
    }
    search_args_t a;
    if (char const* e = prepare_launch(pl, warps, a, s)) return e;
    a.queries = static_cast<uint8_t const*>(d_queries);
    a.query_stride = stride;
    a.nq = (uint32_t)nq;
    a.k = (uint32_t)k;
    a.out_keys = d_keys;
    a.out_dists = d_dists;
    a.out_counts = d_counts;
    a.out_computed = d_computed;
    a.out_visited = d_cycles;
    a.status = status_ptr;
    a.allow_bits = active_allow_bits;
    a.cluster_end_level = active_cluster_end_level;

    if (profile_phases) {
        if (char const* e = phase_cycles.reserve(16)) return e;
        a.phase_cycles = phase_cycles.ptr;
    }
    CU(cudaMemsetAsync(work_counter.ptr, 0, 8, s));
    CU(cudaEventRecord(ev_begin, s));
    CU(search_launch(d, a, blocks, pl.smem_per_block, s));
    CU(cudaEventRecord(ev_end, s));
    kernel_launches += 1;
    if (defer) { /* enqueue only: search_finish synchronises, looks at the status words and retries what overflowed */
        pending.back().args = a;
        return nullptr;
    }
    CU(cudaMemcpyAsync(h_status.ptr, status.ptr, nq * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaEventElapsedTime(&last_kernel_ms, ev_begin, ev_end));
    return retry_overflowed(a, pl.maxed, s);
}

/* every deferred batch: wait, then give the queries whose scratch overflowed their retry */
char const* frozen_index_t::search_finish() {
    char const* first_error = nullptr;
    for (pending_search_t& p : pending) {
        char const* e = cuda_error(cudaStreamSynchronize(p.stream));
        if (!e) e = h_status.reserve(p.args.nq);
        if (!e) e = cuda_error(cudaMemcpy(h_status.ptr, p.status->ptr, (size_t)p.args.nq * 4, cudaMemcpyDeviceToHost));
        if (!e) e = retry_overflowed(p.args, p.maxed, p.stream);
        if (e && !first_error) first_error = e;
        pending_free.push_back(p.status);
    }
    pending.clear();
    if (ev_begin && ev_end && !first_error) cudaEventElapsedTime(&last_kernel_ms, ev_begin, ev_end);
    return first_error;
}

/* scratch overflow (h_status holds the status words of the launch described by `a`): rerun just those queries with 8x
 * larger tables until they fit */
char const* frozen_index_t::retry_overflowed(search_args_t const& a, bool maxed, cudaStream_t s) {
    size_t const nq = a.nq, k = a.k;
    int const wpb = search_warps_per_block();
    std::vector<uint32_t> failed;
    for (size_t i = 0; i < nq; ++i)
        if (h_status.ptr[i] != STATUS_OK) failed.push_back((uint32_t)i);
    uint64_t scale = 1;
    while (!failed.empty()) {
        if (maxed) return "Search scratch overflow that full-size scratch could not fix";
        scale *= 8;
        launch_plan_t rp;
        if (char const* e = plan((uint32_t)k, (uint32_t)std::min<uint64_t>(scale, 1u << 30), rp)) return e;
        maxed = rp.maxed;
        size_t const bytes_per_warp = (size_t)rp.visited_words_per_warp() * 4 + (size_t)rp.heap_spill_cap * 8;
        size_t max_warps = std::max<size_t>(wpb, ((size_t)2 << 30) / bytes_per_warp / wpb * wpb);
        int rblocks = (int)std::min<size_t>({(size_t)rp.blocks, (failed.size() + wpb - 1) / wpb, max_warps / wpb});
        rblocks = std::max(rblocks, 1);
        size_t rwarps = (size_t)rblocks * wpb;
        if (char const* e = retry_list.reserve(failed.size())) return e;
        CU(cudaMemcpyAsync(retry_list.ptr, failed.data(), failed.size() * 4, cudaMemcpyHostToDevice, s));
        search_args_t r;
        if (char const* e = prepare_launch(rp, rwarps, r, s)) return e;
        r.queries = a.queries; r.query_stride = a.query_stride; r.k = a.k;
        r.out_keys = a.out_keys; r.out_dists = a.out_dists; r.out_counts = a.out_counts;
        r.out_computed = a.out_computed; r.out_visited = a.out_visited; r.status = a.status;
        r.allow_bits = a.allow_bits; r.cluster_end_level = a.cluster_end_level; r.phase_cycles = a.phase_cycles;
        r.nq = (uint32_t)failed.size();
        r.query_list = retry_list.ptr;
        CU(cudaMemsetAsync(work_counter.ptr, 0, 8, s));
        CU(search_launch(d, r, rblocks, rp.smem_per_block, s));
        kernel_launches += 1;
        CU(cudaMemcpyAsync(h_status.ptr, a.status, nq * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        std::vector<uint32_t> still;
        for (uint32_t qi : failed)
            if (h_status.ptr[qi] != STATUS_OK) still.push_back(qi);
        failed.swap(still);
    }
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  host-side casts                                                                               */
/* ---------------------------------------------------------------------------------------------- */

namespace {

uint16_t f32_to_f16_rn(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x7FFFFFu;
    int32_t exp = (int32_t)((x >> 23) & 0xFF);
    if (exp == 255) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u | (mant >> 13) : 0));
    exp = exp - 127 + 15;
    if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        mant |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) ++half;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half;
    return (uint16_t)(sign | half);
}

uint16_t f32_to_bf16_rn(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x40u);
    x += 0x7FFFu + ((x >> 16) & 1u);
    return (uint16_t)(x >> 16);
}

} // namespace

char const* cast_queries(uint32_t from, uint32_t to, size_t dims, uint8_t const* src, size_t src_stride, size_t nq,
                         uint8_t* dst, size_t dst_stride) {
    size_t const to_bytes = (dims * bits_per_scalar(to) + 7) / 8;
    if (from == to) {
        for (size_t i = 0; i < nq; ++i) std::memcpy(dst + i * dst_stride, src + i * src_stride, to_bytes);
        return nullptr;
    }
    if (from != SCALAR_F32 && from != SCALAR_F64) return "Only f32/f64 queries can be cast to the index's scalar kind";
    std::vector<float> row(dims);
    for (size_t i = 0; i < nq; ++i) {
        uint8_t const* s = src + i * src_stride;
        uint8_t* o = dst + i * dst_stride;
        if (from == SCALAR_F32) std::memcpy(row.data(), s, dims * 4);
        else
            for (size_t j = 0; j < dims; ++j) { double v; std::memcpy(&v, s + j * 8, 8); row[j] = (float)v; }
        switch (to) {
        case SCALAR_F32: std::memcpy(o, row.data(), dims * 4); break;
        case SCALAR_F16:
            for (size_t j = 0; j < dims; ++j) { uint16_t h = f32_to_f16_rn(row[j]); std::memcpy(o + 2 * j, &h, 2); }
            break;
        case SCALAR_BF16:
            for (size_t j = 0; j < dims; ++j) { uint16_t h = f32_to_bf16_rn(row[j]); std::memcpy(o + 2 * j, &h, 2); }
            break;
        case SCALAR_I8: { /* cast_to_i8_gt, index_plugins.hpp:1172-1191 */
            double magnitude = 0;
            for (size_t j = 0; j < dims; ++j) magnitude += (double)row[j] * (double)row[j];
            magnitude = std::sqrt(magnitude);
            for (size_t j = 0; j < dims; ++j) {
                double v = row[j] * 127.0 / magnitude;
                v = v > 127.0 ? 127.0 : (v < -127.0 ? -127.0 : v);
                reinterpret_cast<int8_t*>(o)[j] = (int8_t)v;
            }
            break;
        }
        case SCALAR_B1: /* cast_to_b1x8_gt, index_plugins.hpp:1139-1158 */
            std::memset(o, 0, to_bytes);
            for (size_t j = 0; j < dims; ++j)
                if (row[j] > 0) o[j / 8] |= (uint8_t)(128 >> (j & 7));
            break;
        default: return "Unsupported scalar kind";
        }
    }
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  batched search on host buffers: H2D + kernel + D2H inside the call                            */
/* ---------------------------------------------------------------------------------------------- */

/* host rows of `query_scalar` -> `queries` on the device in the index's scalar kind, rows zero-padded to vec_stride. The cast
 * (index_dense_gt::search_ casts every query with casts_.from_*, index_dense.hpp:2060-2066) runs on the device. */
char const* frozen_index_t::upload_queries(void const* q, size_t nq, size_t stride, uint32_t query_scalar) {
    size_t const vs = d.vec_stride ? d.vec_stride : 16, bpv = d.bytes_per_vector;
    size_t const src_bytes = (dimensions * bits_per_scalar(query_scalar) + 7) / 8;
    if (!src_bytes) return "Unknown scalar kind!";
    if (stride < src_bytes) { /* single-query callers pass 0; rows can not overlap */
        if (nq != 1 && stride != 0) return "Query stride is smaller than a vector";
        stride = src_bytes;
    }
    if (query_scalar == scalar) {
        if (vs != bpv) CU(cudaMemsetAsync(queries.ptr, 0, nq * vs, stream));
        CU(cudaMemcpy2DAsync(queries.ptr, vs, q, stride, bpv, nq, cudaMemcpyHostToDevice, stream));
        return nullptr;
    }
    if (char const* e = cast_stage.reserve(nq * src_bytes)) return e;
    CU(cudaMemcpy2DAsync(cast_stage.ptr, src_bytes, q, stride, src_bytes, nq, cudaMemcpyHostToDevice, stream));
    return cast_rows_device(cast_stage.ptr, src_bytes, query_scalar, queries.ptr, vs, scalar, dimensions, nq, stream);
}

char const* frozen_index_t::search_host(void const* q, size_t nq, size_t stride, uint32_t query_scalar, size_t k,
                                        uint64_t* keys, size_t keys_stride, float* dists, size_t dists_stride,
                                        size_t* counts, uint64_t* computed_out, uint64_t* cycles_out, size_t* total,
                                        uint64_t const* allowed, size_t allowed_count, bool filtered, int cluster_level) {
    if (total) *total = 0;
    if (nq == 0 || k == 0) return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    if (!loaded || d.n == 0) { /* index_gt::search on an empty index: no matches, no error (index.hpp:3036-3037) */
        if (cluster_level >= 0) return "No clusters to identify";
        for (size_t i = 0; i < nq; ++i) {
            uint64_t* krow = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(keys) + i * keys_stride);
            uint32_t* drow = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(dists) + i * dists_stride);
            for (size_t j = 0; j < k; ++j) { krow[j] = 0; drow[j] = SNAN_BITS; }
            if (counts) counts[i] = 0;
            if (computed_out) computed_out[i] = 0;
            if (cycles_out) cycles_out[i] = 0;
        }
        return nullptr;
    }
    if (char const* e = ensure_context()) return e;
    size_t const vs = d.vec_stride ? d.vec_stride : 16;

    if (char const* e = queries.reserve(nq * vs)) return e;
    if (char const* e = out_keys.reserve(nq * k)) return e;
    if (char const* e = out_dists.reserve(nq * k)) return e;
    if (char const* e = counts_reserve_all(nq)) return e;
    if (char const* e = upload_queries(q, nq, stride, query_scalar)) return e;

    /* filtered search: sort the allowed keys on the host, turn them into a bitmap over slots on the device */
    struct reset_filter_t {
        frozen_index_t* self;
        ~reset_filter_t() { self->active_allow_bits = nullptr; self->active_cluster_end_level = -1; }
    } reset_filter{this};
    /* index_gt::cluster (index.hpp:3115-3116): levels max..`level`, with level 0 treated as level 1 */
    if (cluster_level >= 0) active_cluster_end_level = cluster_level <= 0 ? 0 : cluster_level - 1;
    if (filtered && size) {
        std::vector<uint64_t> sorted(allowed, allowed + allowed_count);
        std::sort(sorted.begin(), sorted.end());
        if (char const* e = allowed_keys.reserve(std::max<size_t>(sorted.size(), 1))) return e;
        if (char const* e = allow_bits.reserve((size + 31) / 32)) return e;
        if (!sorted.empty())
            CU(cudaMemcpyAsync(allowed_keys.ptr, sorted.data(), sorted.size() * 8, cudaMemcpyHostToDevice, stream));
        CU(search_build_allow_bits(d, allowed_keys.ptr, (uint32_t)sorted.size(), allow_bits.ptr, stream));
        CU(cudaStreamSynchronize(stream)); /* `sorted` is pageable host memory */
        active_allow_bits = allow_bits.ptr;
    }

    bool const want_stats = computed_out || cycles_out;
    if (char const* e = search_device(queries.ptr, nq, vs, k, out_keys.ptr, out_dists.ptr, this->counts.ptr,
                                      want_stats ? computed.ptr : nullptr, want_stats ? cycles.ptr : nullptr, stream))
        return e;

    /* results -> host */
    bool const dense = keys_stride == k * 8 && dists_stride == k * 4;
    if (dense) {
        CU(cudaMemcpyAsync(keys, out_keys.ptr, nq * k * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(dists, out_dists.ptr, nq * k * 4, cudaMemcpyDeviceToHost, stream));
    } else {
        CU(cudaMemcpy2DAsync(keys, keys_stride, out_keys.ptr, k * 8, k * 8, nq, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpy2DAsync(dists, dists_stride, out_dists.ptr, k * 4, k * 4, nq, cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaMemcpyAsync(h_counts.ptr, this->counts.ptr, nq * 4, cudaMemcpyDeviceToHost, stream));
    if (want_stats) {
        CU(cudaMemcpyAsync(h_computed.ptr, computed.ptr, nq * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(h_cycles.ptr, cycles.ptr, nq * 4, cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaStreamSynchronize(stream));
    size_t sum = 0;
    for (size_t i = 0; i < nq; ++i) {
        sum += h_counts.ptr[i];
        if (counts) counts[i] = h_counts.ptr[i];
        if (computed_out) computed_out[i] = h_computed.ptr[i];
        if (cycles_out) cycles_out[i] = h_cycles.ptr[i];
    }
    if (total) *total = sum;
    return nullptr;
}

/* One query per call, many calling threads: gather what is waiting into one batch. The first caller to find no leader
 * becomes it, takes every queued request of the same (scalar kind, count) as its own, runs them through search_host as one
 * batch, hands the rows back and repeats until the queue is empty. */
char const* frozen_index_t::search_single(void const* query, uint32_t query_scalar, size_t count, uint64_t* keys, float* dists,
                                          size_t* found) {
    single_request_t mine{query, query_scalar, count, keys, dists, 0, nullptr, false};
    std::unique_lock<std::mutex> lock(gather_mutex);
    gather_queue.push_back(&mine);
    while (!mine.done && gather_leader) gather_cv.wait(lock);
    if (mine.done) { *found = mine.found; return mine.error; }
    gather_leader = true;
    while (!gather_queue.empty()) {
        /* the batch: every waiting request that matches the head's shape */
        std::vector<single_request_t*> batch, rest;
        for (single_request_t* r : gather_queue)
            (r->scalar == gather_queue[0]->scalar && r->count == gather_queue[0]->count ? batch : rest).push_back(r);
        gather_queue.swap(rest);
        lock.unlock();
        size_t const nq = batch.size(), k = batch[0]->count;
        size_t const qbytes = (dimensions * bits_per_scalar(batch[0]->scalar) + 7) / 8;
        char const* e = nullptr;
        if (nq == 1) {
            size_t total = 0;
            e = search_host(batch[0]->query, 1, 0, batch[0]->scalar, k, batch[0]->keys, k * 8, batch[0]->dists, k * 4, nullptr, nullptr,
                            nullptr, &total);
            batch[0]->found = total;
        } else {
            std::vector<uint8_t> q(nq * qbytes);
            std::vector<uint64_t> out_k(nq * k);
            std::vector<float> out_d(nq * k);
            std::vector<size_t> cnt(nq);
            for (size_t i = 0; i < nq; ++i) std::memcpy(q.data() + i * qbytes, batch[i]->query, qbytes);
            e = search_host(q.data(), nq, qbytes, batch[0]->scalar, k, out_k.data(), k * 8, out_d.data(), k * 4, cnt.data(), nullptr,
                            nullptr, nullptr);
            for (size_t i = 0; i < nq && !e; ++i) {
                std::memcpy(batch[i]->keys, out_k.data() + i * k, k * 8);
                std::memcpy(batch[i]->dists, out_d.data() + i * k, k * 4);
                batch[i]->found = cnt[i];
            }
        }
        lock.lock();
        gathered_batches += 1;
        gathered_queries += nq;
        for (single_request_t* r : batch) { r->error = e; r->done = true; }
        gather_cv.notify_all();
    }
    gather_leader = false;
    gather_cv.notify_all(); /* a request that slipped in while the leader was leaving elects a new one */
    *found = mine.found;
    return mine.error;
}

/* ---------------------------------------------------------------------------------------------- */
/*  exact (brute-force) search: host wrappers around exact_kernel.cu                              */
/* ---------------------------------------------------------------------------------------------- */

/* index_gt::search(exact = true) (index.hpp:3047-3051 -> search_exact_ :4251-4268) for a batch of host queries */
char const* frozen_index_t::exact_host(void const* q, size_t nq, size_t stride, uint32_t query_scalar, size_t k, uint64_t* keys,
                                       float* dists, size_t* counts_out) {
    if (nq == 0 || k == 0) return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    if (!loaded || d.n == 0) {
        for (size_t i = 0; i < nq; ++i) {
            for (size_t j = 0; j < k; ++j) { keys[i * k + j] = 0; reinterpret_cast<uint32_t*>(dists)[i * k + j] = SNAN_BITS; }
            if (counts_out) counts_out[i] = 0;
        }
        return nullptr;
    }
    if (char const* e = ensure_context()) return e;
    size_t const vs = d.vec_stride ? d.vec_stride : 16;
    if (char const* e = queries.reserve(nq * vs)) return e;
    if (char const* e = out_keys.reserve(nq * k)) return e;
    if (char const* e = out_dists.reserve(nq * k)) return e;
    if (char const* e = counts_reserve_all(nq)) return e;
    if (char const* e = upload_queries(q, nq, stride, query_scalar)) return e;
    if (char const* e = exact_search_device(d, sm_count, queries.ptr, nq, vs, k, false, false, out_keys.ptr, out_dists.ptr, counts.ptr,
                                            exact_scratch, stream))
        return e;
    kernel_launches += 2;
    CU(cudaMemcpyAsync(keys, out_keys.ptr, nq * k * 8, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(dists, out_dists.ptr, nq * k * 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(h_counts.ptr, counts.ptr, nq * 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (counts_out)
        for (size_t i = 0; i < nq; ++i) counts_out[i] = h_counts.ptr[i];
    return nullptr;
}

/* usearch_exact_search (c/lib.cpp:468-501): many-to-many over raw matrices; keys are dataset row numbers */
char const* exact_search_free(void const* dataset, size_t n, size_t dataset_stride, void const* queries_h, size_t nq, size_t queries_stride,
                              uint32_t scalar, size_t dimensions, uint32_t metric, size_t k, uint64_t* keys, size_t keys_stride,
                              float* distances, size_t distances_stride) {
    if (!search_supported(metric, scalar)) return "This metric / scalar kind has no sm_100a kernel and the backend has no CPU fallback";
    if (!nq || !k) return nullptr;
    if (k > n) return "More neighbours requested than the dataset holds";
    if (n >= 0xFFFFFFFFull) return "Too many entries for 32-bit slots";
    frozen_index_t tmp;
    tmp.device = default_device();
    if (char const* e = tmp.ensure_context()) return e;
    size_t const bpv = (dimensions * bits_per_scalar(scalar) + 7) / 8, vs = (bpv + 15) / 16 * 16;
    device_index_t ix;
    ix.n = (uint32_t)n;
    ix.dims = (uint32_t)dimensions;
    ix.bytes_per_vector = (uint32_t)bpv;
    ix.vec_stride = vs;
    ix.chunks16 = (uint32_t)(vs / 16);
    ix.metric = metric;
    ix.scalar = scalar;
    device_buffer_t<uint8_t> d_vectors, d_queries, scratch;
    device_buffer_t<float> d_norms, d_dists;
    device_buffer_t<uint64_t> d_keys;
    device_buffer_t<uint32_t> d_counts;
    struct release_all_t {
        device_buffer_t<uint8_t>&a, &b, &c; device_buffer_t<float>&d, &e; device_buffer_t<uint64_t>& f; device_buffer_t<uint32_t>& g;
        ~release_all_t() { a.release(); b.release(); c.release(); d.release(); e.release(); f.release(); g.release(); }
    } release_all{d_vectors, d_queries, scratch, d_norms, d_dists, d_keys, d_counts};
    if (char const* e = d_vectors.reserve(n * vs)) return e;
    if (char const* e = d_queries.reserve(nq * vs)) return e;
    if (char const* e = d_keys.reserve(nq * k)) return e;
    if (char const* e = d_dists.reserve(nq * k)) return e;
    if (char const* e = d_counts.reserve(nq)) return e;
    cudaStream_t s = tmp.stream;
    if (vs != bpv) {
        CU(cudaMemsetAsync(d_vectors.ptr, 0, n * vs, s));
        CU(cudaMemsetAsync(d_queries.ptr, 0, nq * vs, s));
    }
    CU(cudaMemcpy2DAsync(d_vectors.ptr, vs, dataset, dataset_stride, bpv, n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpy2DAsync(d_queries.ptr, vs, queries_h, queries_stride, bpv, nq, cudaMemcpyHostToDevice, s));
    ix.vectors = d_vectors.ptr;
    if (search_needs_norms(metric, scalar)) {
        if (char const* e = d_norms.reserve(n)) return e;
        CU(search_compute_norms(ix, d_norms.ptr, s));
        ix.norms = d_norms.ptr;
    }
    if (char const* e = exact_search_device(ix, tmp.sm_count, d_queries.ptr, nq, vs, k, true, true, d_keys.ptr, d_dists.ptr, d_counts.ptr,
                                            scratch, s))
        return e;
    CU(cudaMemcpy2DAsync(keys, keys_stride, d_keys.ptr, k * 8, k * 8, nq, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpy2DAsync(distances, distances_stride, d_dists.ptr, k * 4, k * 4, nq, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return nullptr;
}

/* usearch_distance: both vectors to the device, one warp, the metric struct of the search kernels */
char const* pair_distance_host(void const* a, void const* b, uint32_t scalar, size_t dimensions, uint32_t metric, float* result) {
    frozen_index_t tmp;
    tmp.device = default_device();
    if (char const* e = tmp.ensure_context()) return e;
    size_t const bpv = (dimensions * bits_per_scalar(scalar) + 7) / 8, vs = (bpv + 15) / 16 * 16;
    if (vs > 48 * 1024) return "Vector too long for a single-pair distance";
    device_index_t ix;
    ix.dims = (uint32_t)dimensions;
    ix.bytes_per_vector = (uint32_t)bpv;
    ix.vec_stride = vs;
    ix.chunks16 = (uint32_t)(vs / 16);
    ix.metric = metric;
    ix.scalar = scalar;
    device_buffer_t<uint8_t> pair;
    device_buffer_t<float> out;
    struct release_t { device_buffer_t<uint8_t>& a; device_buffer_t<float>& b; ~release_t() { a.release(); b.release(); } } release{pair, out};
    if (char const* e = pair.reserve(2 * vs)) return e;
    if (char const* e = out.reserve(1)) return e;
    CU(cudaMemsetAsync(pair.ptr, 0, 2 * vs, tmp.stream));
    CU(cudaMemcpyAsync(pair.ptr, a, bpv, cudaMemcpyHostToDevice, tmp.stream));
    CU(cudaMemcpyAsync(pair.ptr + vs, b, bpv, cudaMemcpyHostToDevice, tmp.stream));
    if (char const* e = pair_distance_device(ix, pair.ptr, pair.ptr + vs, out.ptr, tmp.stream)) return e;
    CU(cudaMemcpyAsync(result, out.ptr, 4, cudaMemcpyDeviceToHost, tmp.stream));
    CU(cudaStreamSynchronize(tmp.stream));
    return nullptr;
}

/* ---------------------------------------------------------------------------------------------- */
/*  lookups and edits by key                                                                      */
/* ---------------------------------------------------------------------------------------------- */

/* index_dense_gt::remove (index_dense.hpp:1480-1513): the entry keeps its node and its links, its key becomes the free key
 * (so searches skip it: `deleted_bits`), and the key leaves the lookup table. Slots are not recycled. */
char const* frozen_index_t::remove_key(uint64_t key, size_t* removed) {
    *removed = 0;
    if (!loaded || !size) return nullptr;
    if (char const* e = ensure_context()) return e;
    build_key_map();
    std::vector<uint32_t> slots;
    key_map.for_each(key, [&](uint32_t slot, size_t cell) { slots.push_back(slot); key_map.erase_cell(cell); return true; });
    if (slots.empty()) return nullptr;
    if (!d.deleted_bits) {
        uint32_t* bits = nullptr;
        size_t const words = (capacity + 31) / 32;
        CU(cudaMalloc(&bits, words * 4));
        CU(cudaMemset(bits, 0, words * 4));
        dev_allocs[5] = bits;
        d.deleted_bits = bits;
        hbm_bytes += words * 4;
    }
    for (uint32_t slot : slots) {
        host_keys[slot] = free_key;
        CU(cudaMemcpy(const_cast<uint64_t*>(d.keys) + slot, &free_key, 8, cudaMemcpyHostToDevice));
        uint32_t word = 0;
        CU(cudaMemcpy(&word, d.deleted_bits + (slot >> 5), 4, cudaMemcpyDeviceToHost));
        word |= 1u << (slot & 31);
        CU(cudaMemcpy(const_cast<uint32_t*>(d.deleted_bits) + (slot >> 5), &word, 4, cudaMemcpyHostToDevice));
    }
    count_deleted += slots.size();
    *removed = slots.size();
    return nullptr;
}

/* index_dense_gt::rename (index_dense.hpp:1554-1580): every entry under `from` gets the key `to` */
char const* frozen_index_t::rename_key(uint64_t from, uint64_t to, size_t* renamed) {
    *renamed = 0;
    if (!loaded || !size) return nullptr;
    if (char const* e = ensure_context()) return e;
    if (to == free_key) return "Key is reserved for removed entries";
    build_key_map();
    if (!multi && key_map.contains(to)) return "Renaming impossible, the key is already in use";
    std::vector<uint32_t> slots;
    key_map.for_each(from, [&](uint32_t slot, size_t cell) { slots.push_back(slot); key_map.erase_cell(cell); return true; });
    for (uint32_t slot : slots) {
        host_keys[slot] = to;
        key_map.insert(to, slot);
        CU(cudaMemcpy(const_cast<uint64_t*>(d.keys) + slot, &to, 8, cudaMemcpyHostToDevice));
    }
    *renamed = slots.size();
    return nullptr;
}

namespace {

float half_bits_to_f32(uint16_t h) {
    uint32_t const sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, mant = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (!mant) bits = sign;
        else { /* subnormal: renormalise */
            int e = -1;
            uint32_t m = mant;
            do { ++e; m <<= 1; } while (!(m & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (mant << 13);
    else bits = sign | ((exp + 112u) << 23) | (mant << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

/* one stored vector -> the caller's scalar kind (index_dense_gt::get_ casts with casts_.to_*, index_dense.hpp:2121-2150) */
char const* cast_stored_row(uint32_t from, uint32_t to, size_t dims, uint8_t const* src, uint8_t* dst) {
    size_t const to_bytes = (dims * bits_per_scalar(to) + 7) / 8;
    if (from == to) { std::memcpy(dst, src, to_bytes); return nullptr; }
    std::vector<float> row(dims);
    for (size_t j = 0; j < dims; ++j) {
        switch (from) {
        case SCALAR_F32: std::memcpy(&row[j], src + 4 * j, 4); break;
        case SCALAR_F16: { uint16_t h; std::memcpy(&h, src + 2 * j, 2); row[j] = half_bits_to_f32(h); break; }
        case SCALAR_BF16: { uint16_t h; std::memcpy(&h, src + 2 * j, 2); uint32_t b = (uint32_t)h << 16; std::memcpy(&row[j], &b, 4); break; }
        case SCALAR_I8: row[j] = (float)reinterpret_cast<int8_t const*>(src)[j] / 127.f; break;     /* cast_from_i8_gt */
        case SCALAR_B1: row[j] = (src[j >> 3] & (128u >> (j & 7u))) ? 1.f : 0.f; break;            /* cast_from_b1x8_gt */
        default: return "Unsupported scalar kind";
        }
    }
    if (to == SCALAR_F64) {
        for (size_t j = 0; j < dims; ++j) { double v = row[j]; std::memcpy(dst + 8 * j, &v, 8); }
        return nullptr;
    }
    return cast_queries(SCALAR_F32, to, dims, reinterpret_cast<uint8_t const*>(row.data()), dims * 4, 1, dst, to_bytes);
}

} // namespace

/* index_dense_gt::get (index_dense.hpp:781-786 -> get_ :2121-2150): up to `max_count` vectors stored under `key` */
char const* frozen_index_t::get_vectors(uint64_t key, size_t max_count, void* out, uint32_t out_scalar, size_t* found) {
    *found = 0;
    if (!loaded || !size || !max_count) return nullptr;
    if (char const* e = ensure_context()) return e;
    size_t const out_bytes = (dimensions * bits_per_scalar(out_scalar) + 7) / 8;
    if (!out_bytes) return "Unknown scalar kind!";
    build_key_map();
    std::vector<uint32_t> slots;
    key_map.for_each(key, [&](uint32_t slot, size_t) { slots.push_back(slot); return slots.size() < max_count; });
    std::sort(slots.begin(), slots.end()); /* insertion order */
    std::vector<uint8_t> row(d.vec_stride);
    for (size_t i = 0; i < slots.size(); ++i) {
        CU(cudaMemcpy(row.data(), d.vectors + (size_t)slots[i] * d.vec_stride, d.bytes_per_vector, cudaMemcpyDeviceToHost));
        if (char const* e = cast_stored_row(scalar, out_scalar, dimensions, row.data(), static_cast<uint8_t*>(out) + i * out_bytes)) return e;
    }
    *found = slots.size();
    return nullptr;
}

} // namespace usearch_b200
