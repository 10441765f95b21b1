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
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "device_index.h"
#include "key_map.h"

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
int search_max_warps_per_sm(device_index_t const& ix);
bool search_single_stage_set(device_index_t const& ix);
bool search_needs_norms(uint32_t metric, uint32_t scalar);
cudaError_t search_compute_norms(device_index_t const& ix, float* norms, cudaStream_t stream);
cudaError_t search_fill_empty(uint64_t* keys, float* dists, uint32_t* counts, uint32_t* computed, uint32_t* visited, size_t nq,
                              size_t k, cudaStream_t stream);
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
    uint32_t off_bars = 0, off_stage = 0, stage_stride = 0, stage_sets = 1;
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

/* device scratch of the batched builder (builder.cu) */
struct build_scratch_t {
    device_buffer_t<uint32_t> task_slot, cand_slots, cand_counts, pair_idx, pair_idx_sorted, heads, counters;
    device_buffer_t<uint8_t> task_level, sort_temp;
    device_buffer_t<float> cand_dists, pair_dists;
    device_buffer_t<uint64_t> pair_keys, pair_keys_sorted;
    size_t iota_count = 0;
    void release() {
        task_slot.release(); cand_slots.release(); cand_counts.release(); pair_idx.release(); pair_idx_sorted.release();
        heads.release(); counters.release(); task_level.release(); sort_temp.release(); cand_dists.release();
        pair_dists.release(); pair_keys.release(); pair_keys_sorted.release();
        iota_count = 0;
    }
};

struct shard_group_t; /* shards.cu */

struct frozen_index_t {
    /* configuration (usearch_init_options_t) */
    uint32_t metric = 0, scalar = 0; /* reference char codes */
    size_t dimensions = 0, connectivity = 0, connectivity_base = 0;
    size_t expansion_add = 128, expansion_search = 64; /* index.hpp:1340-1350 defaults */
    bool multi = false;
    uint64_t free_key = UINT64_MAX; /* index_dense.hpp:513 */

    /* population */
    size_t size = 0, count_deleted = 0;
    std::vector<int16_t> levels;     /* kept on the host: re-serialisation and the builder's work lists */
    std::vector<uint64_t> host_keys; /* host copy of `keys` (slot -> key): lookups by key never touch the device */
    key_map_t key_map;
    size_t capacity = 0;                      /* slots the HBM arrays have room for */
    size_t upper_capacity = 0, upper_rows = 0; /* rows of `upper`: allocated / in use */
    uint64_t level_seed = 0;
    bool configured() const { return metric && scalar && dimensions && connectivity; }
    void build_key_map() { if (!key_map.built) key_map.rebuild(host_keys, free_key, capacity); }

    /* device */
    int device = 0;
    device_index_t d;
    size_t hbm_bytes = 0;
    bool loaded = false;
    void* dev_allocs[8] = {nullptr};

    /* tuning knobs of the search launch: environment at construction (USEARCH_B200_STAGE_SETS, _WARPS_PER_SM),
     * changeable per handle with usearch_b200_tune (bench sweeps, tests) */
    struct tune_t {
        int stage_sets = env_int("USEARCH_B200_STAGE_SETS", 0);     /* 0 = planned, 1 or 2 = forced */
        int warps_per_sm = env_int("USEARCH_B200_WARPS_PER_SM", 0); /* 0 = as many as fit, else an upper bound */
        static int env_int(char const* name, int fallback) {
            char const* v = std::getenv(name);
            return v ? std::atoi(v) : fallback;
        }
    } tune;

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

    /* mutation (builder.cu): GPU-assisted add, capacity, tombstones */
    build_scratch_t build;
    device_buffer_t<uint8_t> cast_stage; /* raw caller rows awaiting a scalar cast on the device */
    char const* reserve_slots(size_t slots);
    char const* reserve_upper_rows(size_t rows);
    int16_t draw_level(size_t slot) const;
    char const* add_many(uint64_t const* keys, void const* vectors, size_t count, size_t stride, uint32_t scalar_kind, bool on_device);
    char const* link_batch(size_t first, size_t count);
    char const* remove_key(uint64_t key, size_t* removed);
    char const* rename_key(uint64_t from, uint64_t to, size_t* renamed);
    char const* get_vectors(uint64_t key, size_t max_count, void* out, uint32_t out_scalar, size_t* found);

    /* sharded search (shards.cu): this handle is shard `rank` of `world`, one process per GPU */
    shard_group_t* shards = nullptr;
    char const* join_shards(int rank, int world, void const* unique_id128);
    void leave_shards();
    char const* sharded_search_device(void const* d_queries, size_t nq, size_t stride, size_t k, uint64_t* d_keys, float* d_dists,
                                      uint32_t* d_counts, uint32_t* d_computed, uint32_t* d_cycles, cudaStream_t stream);
    char const* sharded_search_host(void const* queries, size_t nq, size_t stride, uint32_t query_scalar, size_t k, uint64_t* keys,
                                    float* dists, size_t* counts);

    /* searches */
    char const* plan(uint32_t k, uint32_t visited_cap_override, launch_plan_t& plan, uint32_t ef_override = 0) const;
    char const* prepare_launch(launch_plan_t const& pl, size_t warps, search_args_t& a, cudaStream_t s);
    char const* search_device(void const* d_queries, size_t nq, size_t stride, size_t k, uint64_t* d_keys, float* d_dists,
                              uint32_t* d_counts, uint32_t* d_computed, uint32_t* d_cycles, cudaStream_t stream, bool defer = false);
    /* deferred launches (usearch_b200_search_many_enqueue): status words and arguments kept until search_finish */
    struct pending_search_t { device_buffer_t<uint32_t>* status; search_args_t args; bool maxed; cudaStream_t stream; };
    std::vector<pending_search_t> pending;
    std::vector<device_buffer_t<uint32_t>*> pending_free;
    char const* search_finish();
    char const* retry_overflowed(search_args_t const& a, bool maxed, cudaStream_t stream);
    /* usearch_search from many host threads: callers that arrive while a launch is in flight are gathered and served by
     * ONE launch (a leader runs the batch, the others wait for their rows) — the reference serves them from distinct
     * thread contexts in parallel (index.hpp:3033-3039, index_dense.hpp:1984-2000) */
    struct single_request_t {
        void const* query; uint32_t scalar; size_t count; uint64_t* keys; float* dists; size_t found; char const* error; bool done;
    };
    std::mutex gather_mutex;
    std::condition_variable gather_cv;
    std::vector<single_request_t*> gather_queue;
    bool gather_leader = false;
    uint64_t gathered_batches = 0, gathered_queries = 0;
    char const* search_single(void const* query, uint32_t query_scalar, size_t count, uint64_t* keys, float* dists, size_t* found);
    char const* upload_queries(void const* queries, size_t nq, size_t stride, uint32_t query_scalar);
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

/* shards.cu */
char const* shards_unique_id(void* out128);
size_t shards_payload_bytes(size_t nq, size_t k);
char const* shards_merge_host(void const* payloads, int world, size_t nq, size_t k, uint64_t* keys, float* dists, uint32_t* counts);

/* builder.cu: scalar casts and single-pair distances on the device */
char const* cast_rows_device(uint8_t const* src, size_t src_stride, uint32_t from, uint8_t* dst, size_t dst_stride, uint32_t to,
                             size_t dims, size_t rows, cudaStream_t stream);
char const* pair_distance_device(device_index_t const& shape, uint8_t const* d_a, uint8_t const* d_b, float* d_out, cudaStream_t stream);

char const* pair_distance_host(void const* a, void const* b, uint32_t scalar, size_t dimensions, uint32_t metric, float* result);
int default_device(); /* USEARCH_B200_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0 */

/* host-side query casts (index_plugins.hpp:1105-1224) */
char const* cast_queries(uint32_t from_scalar, uint32_t to_scalar, size_t dims, uint8_t const* src, size_t src_stride, size_t nq,
                         uint8_t* dst, size_t dst_stride);
size_t bits_per_scalar(uint32_t scalar);

} // namespace usearch_b200
