/*
 *  frozen_index.h — host-side owner of one HNSW index frozen into HBM, plus the scratch, streams
 *  and staging buffers its searches use. This is the object behind the opaque `usearch_index_t`
 *  of include/usearch_b200.h; it plays the role `index_dense_gt` plays behind the reference's
 *  handle (c/lib.cpp:136-182) for the search path only.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "device_index.h"

namespace usearch_b200 {

/* kernel entry points (search_kernel.cu) */
cudaError_t search_launch(device_index_t const& ix, search_args_t const& a, int blocks, size_t smem, cudaStream_t stream);
cudaError_t search_occupancy(device_index_t const& ix, int* blocks_per_sm, size_t smem);
bool search_supported(uint32_t metric, uint32_t scalar);
int search_warps_per_block();
bool search_is_staged(device_index_t const& ix);
int search_stage_slots(device_index_t const& ix);
int search_lanes_per_vector(device_index_t const& ix);
uint32_t search_stage_pad(device_index_t const& ix);
bool search_needs_norms(uint32_t metric, uint32_t scalar);
cudaError_t search_compute_norms(device_index_t const& ix, float* norms, cudaStream_t stream);
cudaError_t search_build_allow_bits(device_index_t const& ix, uint64_t const* allowed_sorted, uint32_t m, uint32_t* bits,
                                    cudaStream_t stream);

/* `visits` is a per-warp bitmap whenever all the bitmaps fit this budget (else an open-addressing table);
 * bitmaps above BITMAP_WIPE_MAX_SLOTS are cleaned through a log of the bits each query set */
constexpr uint64_t BITMAP_SCRATCH_BUDGET = 12ull << 30;
constexpr uint64_t BITMAP_WIPE_MAX_SLOTS = 1ull << 20;

struct launch_plan_t {
    uint32_t ef = 0;
    uint32_t visited_cap = 0, visited_bitmap_words = 0, visit_log_cap = 0, heap_spill_cap = 0, heap_smem_cap = 0;
    bool maxed = false; /* growing the scratch any further cannot help */
    size_t visited_words_per_warp() const { return visited_bitmap_words ? visited_bitmap_words : visited_cap; }
    uint32_t smem_per_warp = 0, off_top_d = 0, off_top_s = 0, off_cand_s = 0, off_cand_d = 0, off_heap = 0;
    uint32_t off_bars = 0, off_stage = 0, stage_stride = 0, stage_sets = 1, stage_segments = 1, stage_seg_chunks = 0;
    int blocks = 0;
    uint32_t warps_per_sm_target = 0;
    size_t smem_per_block = 0;
    size_t warps() const { return (size_t)blocks * (size_t)search_warps_per_block(); }
};

template <typename T> struct device_buffer_t {
    T* ptr = nullptr;
    size_t capacity = 0; /* elements */
    char const* reserve(size_t n) {
        if (n <= capacity) return nullptr;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        capacity = 0;
        if (cudaMalloc(&ptr, n * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            return "Out of GPU memory!";
        }
        capacity = n;
        return nullptr;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        capacity = 0;
    }
};

template <typename T> struct pinned_buffer_t {
    T* ptr = nullptr;
    size_t capacity = 0;
    char const* reserve(size_t n) {
        if (n <= capacity) return nullptr;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        capacity = 0;
        if (cudaHostAlloc(&ptr, n * sizeof(T), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return "Out of pinned host memory!";
        }
        capacity = n;
        return nullptr;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        capacity = 0;
    }
};

struct frozen_index_t {
    /* configuration (usearch_init_options_t) */
    uint32_t metric = 0, scalar = 0; /* reference char codes */
    size_t dimensions = 0, connectivity = 0, connectivity_base = 0;
    size_t expansion_add = 128, expansion_search = 64; /* index.hpp:1340-1350 defaults */
    bool multi = false;
    uint64_t free_key = UINT64_MAX; /* index_dense.hpp:513 */

    /* population */
    size_t size = 0, count_deleted = 0;
    std::vector<int16_t> levels; /* kept on the host: only needed to re-serialise */

    /* device */
    int device = 0;
    device_index_t d;
    size_t hbm_bytes = 0;
    bool loaded = false;
    void* dev_allocs[8] = {nullptr};

    /* per-handle execution context */
    std::mutex mutex;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    device_buffer_t<uint32_t> visited, visit_log, work_counter, status, counts, computed, cycles, retry_list;
    size_t visited_zeroed_words = 0; /* the first this-many words of `visited` are known to be zero (logged bitmaps) */
    device_buffer_t<cand_t> heap_spill;
    device_buffer_t<uint8_t> queries;
    device_buffer_t<uint64_t> allowed_keys; /* filtered search: sorted allowed keys and the bitmap built from them */
    device_buffer_t<uint32_t> allow_bits;
    uint32_t const* active_allow_bits = nullptr; /* set for the duration of one filtered call */
    int active_cluster_end_level = -1;           /* set for the duration of one cluster() call */
    device_buffer_t<uint64_t> out_keys;
    device_buffer_t<float> out_dists;
    pinned_buffer_t<uint8_t> h_queries;
    pinned_buffer_t<uint64_t> h_keys;
    pinned_buffer_t<float> h_dists;
    pinned_buffer_t<uint32_t> h_counts, h_computed, h_cycles, h_status;
    device_buffer_t<unsigned long long> phase_cycles; /* introspection, enabled by usearch_b200_profile_phases */
    bool profile_phases = false;
    uint64_t kernel_launches = 0;
    float last_kernel_ms = 0.f;
    int sm_count = 0;

    ~frozen_index_t();
    void release_device();
    char const* ensure_context();
    char const* counts_reserve_all(size_t nq);

    /* v2 blob -> HBM (index_dense.hpp:1084-1188, index.hpp:3322-3382) */
    char const* load_blob(uint8_t const* blob, size_t length);
    size_t serialized_length() const;
    char const* save_blob(uint8_t* out, size_t length) const;

    /* searches */
    char const* plan(uint32_t k, uint32_t visited_cap_override, launch_plan_t& plan) const;
    char const* search_device(void const* d_queries, size_t nq, size_t stride, size_t k, uint64_t* d_keys, float* d_dists,
                              uint32_t* d_counts, uint32_t* d_computed, uint32_t* d_cycles, cudaStream_t stream);
    device_buffer_t<uint8_t> exact_scratch;
    char const* exact_host(void const* queries, size_t nq, size_t stride, uint32_t query_scalar, size_t k, uint64_t* keys,
                           float* dists, size_t* counts);
    char const* search_host(void const* queries, size_t nq, size_t stride, uint32_t query_scalar, size_t k, uint64_t* keys,
                            size_t keys_stride, float* dists, size_t dists_stride, size_t* counts, uint64_t* computed,
                            uint64_t* cycles, size_t* total, uint64_t const* allowed = nullptr, size_t allowed_count = 0,
                            bool filtered = false, int cluster_level = -1);
};

/* exact_kernel.cu */
char const* exact_search_device(device_index_t const& ix, int sm_count, void const* d_queries, size_t nq, size_t query_stride, size_t k,
                                bool swap, bool slots_as_keys, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                                device_buffer_t<uint8_t>& scratch, cudaStream_t stream);
char const* exact_search_free(void const* dataset, size_t dataset_count, size_t dataset_stride, void const* queries,
                              size_t queries_count, size_t queries_stride, uint32_t scalar, size_t dimensions, uint32_t metric,
                              size_t count, uint64_t* keys, size_t keys_stride, float* distances, size_t distances_stride);

/* host-side query casts (index_plugins.hpp:1105-1224) */
char const* cast_queries(uint32_t from_scalar, uint32_t to_scalar, size_t dims, uint8_t const* src, size_t src_stride, size_t nq,
                         uint8_t* dst, size_t dst_stride);
size_t bits_per_scalar(uint32_t scalar);

} // namespace usearch_b200
