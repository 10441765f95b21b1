/*
 *  usearch_b200.h — C ABI of the B200-native batched HNSW search backend.
 *
 *  The library (usearch_b200/libusearch_b200.so) is a drop-in for the SEARCH path of the
 *  reference's C ABI. Every `usearch_*` symbol below has the name, argument order, argument
 *  meaning and error convention of the declaration it replaces in the reference header
 *  `c/usearch.h` (cited per function as usearch.h:LINE, implementation c/lib.cpp:LINE), so that
 *  Go (golang/lib.go:29-33), C# (NativeMethods.cs:16) and C callers bind it unchanged.
 *
 *  Division of labour: the graph is BUILT by the host path (the reference itself,
 *  or any writer of the v2 `.usearch` format), serialised, and handed to this library through
 *  `usearch_load[_buffer]` / `usearch_view[_buffer]`, which freeze it into a flat SoA layout in
 *  HBM. From then on every `usearch_search` / `usearch_search_many` runs on the GPU. Mutating
 *  entry points are exported so that existing bindings link, and report
 *  "Index is frozen in GPU memory ..." through `error` (the reference's own convention for an
 *  immutable `view`, index.hpp:2787-2788).
 *
 *  Error convention (usearch.h:24-28): `*error` receives a pointer to a static, NUL-terminated
 *  message that must not be freed; it is left untouched on success. Nothing throws across the ABI.
 *
 *  `usearch_b200_*` symbols are additive: the batch entry the reference lacks (SURVEY.md
 *  finding 2; python/lib.cpp:286-308 loops single-query calls on a thread pool instead), a
 *  device-pointer variant for callers that already hold queries in HBM, and introspection.
 */
#ifndef USEARCH_B200_H
#define USEARCH_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types: identical to usearch.h:20-110 ------------------------------------------------ */

typedef void* usearch_index_t;
typedef uint64_t usearch_key_t;
typedef float usearch_distance_t;
typedef char const* usearch_error_t;
typedef usearch_distance_t (*usearch_metric_t)(void const*, void const*);

typedef enum usearch_metric_kind_t { /* usearch.h:40-52 */
    usearch_metric_unknown_k = 0,
    usearch_metric_cos_k = 1,
    usearch_metric_ip_k = 2,
    usearch_metric_l2sq_k = 3,
    usearch_metric_haversine_k = 4,
    usearch_metric_divergence_k = 5,
    usearch_metric_pearson_k = 6,
    usearch_metric_jaccard_k = 7,
    usearch_metric_hamming_k = 8,
    usearch_metric_tanimoto_k = 9,
    usearch_metric_sorensen_k = 10,
} usearch_metric_kind_t;

typedef enum usearch_scalar_kind_t { /* usearch.h:54-62 */
    usearch_scalar_unknown_k = 0,
    usearch_scalar_f32_k = 1,
    usearch_scalar_f64_k = 2,
    usearch_scalar_f16_k = 3,
    usearch_scalar_i8_k = 4,
    usearch_scalar_b1_k = 5,
    usearch_scalar_bf16_k = 6,
} usearch_scalar_kind_t;

typedef struct usearch_init_options_t { /* usearch.h:64-110, same field order */
    usearch_metric_kind_t metric_kind;
    usearch_metric_t metric; /* custom host callbacks cannot run on the device: must be NULL */
    usearch_scalar_kind_t quantization;
    size_t dimensions;
    size_t connectivity;
    size_t expansion_add;
    size_t expansion_search;
    bool multi;
} usearch_init_options_t;

/* ---- lifecycle & introspection ----------------------------------------------------------- */

char const* usearch_version(void);                                                      /* usearch.h:116 */
usearch_index_t usearch_init(usearch_init_options_t* options, usearch_error_t* error);  /* usearch.h:124; NULL options = empty index awaiting load (c/lib.cpp:142-147) */
void usearch_free(usearch_index_t index, usearch_error_t* error);                       /* usearch.h:131 */
size_t usearch_memory_usage(usearch_index_t index, usearch_error_t* error);             /* usearch.h:139; bytes of HBM held */
char const* usearch_hardware_acceleration(usearch_index_t index, usearch_error_t* error); /* usearch.h:147; "sm_100a" */
size_t usearch_serialized_length(usearch_index_t index, usearch_error_t* error);        /* usearch.h:154 */

/* ---- the hand-off: v2 `.usearch` blob → HBM (index_dense.hpp:1084-1188, index.hpp:3322-3382) */

void usearch_save(usearch_index_t index, char const* path, usearch_error_t* error);     /* usearch.h:162 */
void usearch_load(usearch_index_t index, char const* path, usearch_error_t* error);     /* usearch.h:170 */
void usearch_view(usearch_index_t index, char const* path, usearch_error_t* error);     /* usearch.h:178; same as load: the device copy never aliases the file */
void usearch_metadata(char const* path, usearch_init_options_t* options, usearch_error_t* error); /* usearch.h:186 */
void usearch_save_buffer(usearch_index_t index, void* buffer, size_t length, usearch_error_t* error);        /* usearch.h:195 */
void usearch_load_buffer(usearch_index_t index, void const* buffer, size_t length, usearch_error_t* error);  /* usearch.h:204 */
void usearch_view_buffer(usearch_index_t index, void const* buffer, size_t length, usearch_error_t* error);  /* usearch.h:214 */
void usearch_metadata_buffer(void const* buffer, size_t length, usearch_init_options_t* options, usearch_error_t* error); /* usearch.h:223 */

size_t usearch_size(usearch_index_t index, usearch_error_t* error);          /* usearch.h:231 */
size_t usearch_capacity(usearch_index_t index, usearch_error_t* error);      /* usearch.h:238 */
size_t usearch_dimensions(usearch_index_t index, usearch_error_t* error);    /* usearch.h:245 */
size_t usearch_connectivity(usearch_index_t index, usearch_error_t* error);  /* usearch.h:252 */
void usearch_reserve(usearch_index_t index, size_t capacity, usearch_error_t* error); /* usearch.h:260, c/lib.cpp:365-370: room for `capacity` members in HBM */
size_t usearch_expansion_add(usearch_index_t index, usearch_error_t* error);          /* usearch.h:268 */
size_t usearch_expansion_search(usearch_index_t index, usearch_error_t* error);       /* usearch.h:276 */
void usearch_change_expansion_add(usearch_index_t index, size_t expansion, usearch_error_t* error);    /* usearch.h:284 */
void usearch_change_expansion_search(usearch_index_t index, size_t expansion, usearch_error_t* error); /* usearch.h:292 */
void usearch_change_threads_add(usearch_index_t index, size_t threads, usearch_error_t* error);        /* usearch.h:300; accepted, ignored */
void usearch_change_threads_search(usearch_index_t index, size_t threads, usearch_error_t* error);     /* usearch.h:308; accepted, ignored */
void usearch_change_metric_kind(usearch_index_t index, usearch_metric_kind_t kind, usearch_error_t* error); /* usearch.h:316 */
void usearch_change_metric(usearch_index_t index, usearch_metric_t metric, void* state, usearch_metric_kind_t kind, usearch_error_t* error); /* usearch.h:327; host callbacks rejected */

/* ---- the hot path ------------------------------------------------------------------------ */

/* usearch.h:371-374, c/lib.cpp:398-411. Returns the number of matches; ALWAYS writes `count`
 * output slots, padding with key 0 / signalling-NaN distance (index.hpp:2707-2722). The query
 * may be of any scalar kind; it is cast to the index's kind with the reference's rules
 * (index_plugins.hpp:1105-1224). */
size_t usearch_search(usearch_index_t index, void const* query_vector, usearch_scalar_kind_t query_kind, size_t count,
                      usearch_key_t* keys, usearch_distance_t* distances, usearch_error_t* error);

/* usearch.h:391-395. Host callbacks cannot run on the device: reports an error unless `filter` is NULL. */
size_t usearch_filtered_search(usearch_index_t index, void const* query_vector, usearch_scalar_kind_t query_kind,
                               size_t count, int (*filter)(usearch_key_t key, void* filter_state), void* filter_state,
                               usearch_key_t* keys, usearch_distance_t* distances, usearch_error_t* error);

/* NEW (additive). One call = one batch = one persistent-kernel launch. Strides are in BYTES,
 * in the style of usearch_exact_search (usearch.h:467-474). `counts` receives the per-query
 * number of matches. Replaces the thread-pool loop in python/lib.cpp:261-319 (`search_typed`)
 * and cpp/bench.cpp:352-377. Buffers are HOST memory; copies are part of the call. Returns the
 * sum of counts. */
size_t usearch_search_many(usearch_index_t index, void const* queries, size_t queries_count, size_t queries_stride,
                           usearch_scalar_kind_t query_kind, size_t count,              //
                           usearch_key_t* keys, size_t keys_stride,                     //
                           usearch_distance_t* distances, size_t distances_stride,      //
                           size_t* counts, usearch_error_t* error);

/* ---- mutation and lookups by key ---------------------------------------------------------------- */

/* usearch.h:338, c/lib.cpp:378-386 -> index_gt::add (index.hpp:2780-2880). The member is linked into the graph on the GPU
 * by the batched builder (csrc/builder.cu) — a batch of one here; usearch_b200_add_many below is the throughput entry.
 * Capacity grows on demand (the reference's C layer reports "Reserve capacity ahead of insertions!" instead). Removed
 * entries keep their slot (tombstone, skipped by searches); slots are not recycled. */
void usearch_add(usearch_index_t index, usearch_key_t key, void const* vector, usearch_scalar_kind_t vector_kind, usearch_error_t* error); /* usearch.h:338 */
bool usearch_contains(usearch_index_t index, usearch_key_t key, usearch_error_t* error);   /* usearch.h:349 */
size_t usearch_count(usearch_index_t index, usearch_key_t key, usearch_error_t* error);    /* usearch.h:358 */
size_t usearch_get(usearch_index_t index, usearch_key_t key, size_t count, void* vector, usearch_scalar_kind_t vector_kind, usearch_error_t* error); /* usearch.h:407 */
size_t usearch_remove(usearch_index_t index, usearch_key_t key, usearch_error_t* error);   /* usearch.h:418 */
size_t usearch_rename(usearch_index_t index, usearch_key_t from, usearch_key_t to, usearch_error_t* error); /* usearch.h:428 */
usearch_distance_t usearch_distance(void const* vector_first, void const* vector_second, usearch_scalar_kind_t scalar_kind,
                                    size_t dimensions, usearch_metric_kind_t metric_kind, usearch_error_t* error); /* usearch.h:441 */
void usearch_exact_search(void const* dataset, size_t dataset_size, size_t dataset_stride, void const* queries,
                          size_t queries_size, size_t queries_stride, usearch_scalar_kind_t scalar_kind, size_t dimensions,
                          usearch_metric_kind_t metric_kind, size_t count, size_t threads, usearch_key_t* keys,
                          size_t keys_stride, usearch_distance_t* distances, size_t distances_stride,
                          usearch_error_t* error); /* usearch.h:467; runs on the GPU, `threads` ignored; count > 256 needs vectors that fit the tiled stage (about 3.5 KB) */
void usearch_clear(usearch_index_t index, usearch_error_t* error); /* usearch.h:481 */

/* NEW (additive). GPU-assisted construction: `count` keys and vectors in one call. The vectors (any
 * supported scalar kind, rows `vectors_stride` bytes apart, 0 = dense) are copied into HBM, cast on the device if needed, and
 * linked batch by batch: one INSERT-mode launch of the search kernel per batch produces every member's candidates on every
 * level (search_to_insert_, index.hpp:4010-4079), two more kernels select forward and reverse links with the reference's
 * heuristic (refine_, index.hpp:4276-4318). Replaces the thread-pool loop of python/lib.cpp:171-258 (`add_typed_to_index`). */
void usearch_b200_add_many(usearch_index_t index, usearch_key_t const* keys, void const* vectors, size_t count,
                           size_t vectors_stride, usearch_scalar_kind_t vector_kind, usearch_error_t* error);
/* The same with `keys` and `vectors` in DEVICE memory on the index's GPU. */
void usearch_b200_add_many_device(usearch_index_t index, usearch_key_t const* keys, void const* vectors, size_t count,
                                  size_t vectors_stride, usearch_scalar_kind_t vector_kind, usearch_error_t* error);

/* ---- additive: sharded search, one process per GPU ------------------------------------------------ */

/* The reference's `Indexes` (python/lib.cpp:74-107, :321-402) searches every query in every shard and merges by distance. Here
 * each process holds one shard on its GPU. Rank 0 obtains a 128-byte id (an ncclUniqueId), the host side hands it to every
 * rank over its own control plane (torch.distributed, MPI, a file), every rank joins. A sharded search = this shard's batched
 * search + ONE NCCL all-gather of the packed per-shard rows + a k-way merge kernel ordered by (distance, shard, position);
 * every rank receives the merged rows. Collective: all ranks must call with the same queries and count. */
void usearch_b200_shards_unique_id(void* unique_id128, usearch_error_t* error);
void usearch_b200_shards_join(usearch_index_t index, int rank, int world, void const* unique_id128, usearch_error_t* error);
size_t usearch_b200_sharded_search_many(usearch_index_t index, void const* queries, size_t queries_count, size_t queries_stride,
                                        usearch_scalar_kind_t query_kind, size_t count, usearch_key_t* keys,
                                        usearch_distance_t* distances, size_t* counts, usearch_error_t* error);
/* device pointers, queries already in the index's scalar kind (see usearch_b200_search_many_device). With a `cuda_stream` the
 * all-gather and the merge are only ENQUEUED on it (the outputs are ready in stream order); with NULL the call uses the handle's
 * own stream and returns when the merged rows are complete. */
void usearch_b200_sharded_search_many_device(usearch_index_t index, void const* queries, size_t queries_count,
                                             size_t queries_stride, size_t count, usearch_key_t* keys,
                                             usearch_distance_t* distances, uint32_t* counts, uint32_t* computed_distances,
                                             uint32_t* visited_members, void* cuda_stream, usearch_error_t* error);
/* The merge on its own (search_result_t::merge_into, index.hpp:2650-2670, made deterministic), for callers that hold several
 * shards in one process: `payloads` = `world` blocks of usearch_b200_shards_payload_bytes(queries_count, count) bytes in HOST
 * memory, each `keys u64[nq*count] | distances f32[nq*count] | counts u32[nq]` (16-byte padded). */
size_t usearch_b200_shards_payload_bytes(size_t queries_count, size_t count);
void usearch_b200_merge_topk(void const* payloads, int world, size_t queries_count, size_t count, usearch_key_t* keys,
                             usearch_distance_t* distances, uint32_t* counts, usearch_error_t* error);

/* ---- additive, device-resident variants ---------------------------------------------------- */

/* All pointers are DEVICE pointers on the index's GPU; `queries` must already be in the index's
 * scalar kind, rows `queries_stride` bytes apart (a multiple of 16 or equal to bytes-per-vector).
 * `keys`/`distances` are dense [queries_count x count]; `counts`, `computed_distances` and
 * `visited_members` (the reference's per-query counters, index.hpp:2605-2609; may be NULL) are
 * uint32 [queries_count]. `cuda_stream` is a cudaStream_t (NULL = the handle's own stream). Nothing is copied; the call
 * returns when the batch is complete (it waits for the kernel to read the per-query status words and retries scratch
 * overflows). For launches that do not wait, see usearch_b200_search_many_enqueue / _finish below. */
void usearch_b200_search_many_device(usearch_index_t index, void const* queries, size_t queries_count,
                                     size_t queries_stride, size_t count, usearch_key_t* keys,
                                     usearch_distance_t* distances, uint32_t* counts, uint32_t* computed_distances,
                                     uint32_t* visited_members, void* cuda_stream, usearch_error_t* error);

/* The asynchronous pair. `enqueue` = the same arguments as usearch_b200_search_many_device, but it ONLY enqueues the
 * kernel on `cuda_stream` and returns; any number of batches may be in flight. `finish` waits for them, inspects the
 * per-query status words and re-runs, with larger scratch, the rare queries whose scratch overflowed; the outputs of
 * every enqueued batch are final when it returns. */
void usearch_b200_search_many_enqueue(usearch_index_t index, void const* queries, size_t queries_count,
                                      size_t queries_stride, size_t count, usearch_key_t* keys,
                                      usearch_distance_t* distances, uint32_t* counts, uint32_t* computed_distances,
                                      uint32_t* visited_members, void* cuda_stream, usearch_error_t* error);
void usearch_b200_search_many_finish(usearch_index_t index, usearch_error_t* error);

/* Like usearch_search_many (host buffers) but also returns the reference's two counters. */
size_t usearch_b200_search_many_stats(usearch_index_t index, void const* queries, size_t queries_count,
                                      size_t queries_stride, usearch_scalar_kind_t query_kind, size_t count,
                                      usearch_key_t* keys, usearch_distance_t* distances, size_t* counts,
                                      uint64_t* computed_distances, uint64_t* visited_members, usearch_error_t* error);

/* Introspection for tests / bench: CUDA device ordinal, kernel launches issued so far by this
 * handle, duration in milliseconds of the most recent search kernel (CUDA events on its stream). */
/* Device-side counterpart of usearch_filtered_search (usearch.h:391-395) for the one predicate family that
 * can run on a GPU: "the key is in this set". `allowed_keys` (host memory, any order, may be empty) is turned
 * into a bitmap over slots; the predicate is applied where the reference applies its callback
 * (index_dense.hpp:2078-2083, index.hpp:4201/4236): rejected members are still traversed, never returned. */
size_t usearch_b200_filtered_search_many(usearch_index_t index, void const* queries, size_t queries_count,
                                         size_t queries_stride, usearch_scalar_kind_t query_kind, size_t count,
                                         usearch_key_t const* allowed_keys, size_t allowed_count, usearch_key_t* keys,
                                         usearch_distance_t* distances, size_t* counts, uint64_t* computed_distances,
                                         uint64_t* visited_members, usearch_error_t* error);

/* `index_dense_gt::cluster(vector, level)` (index_dense.hpp:788-793, index.hpp:3092-3125) for a batch: the closest
 * member on graph level `level` for every query (greedy descent only). Outputs hold one entry per query;
 * computed_distances / visited_members may be NULL. */
void usearch_b200_cluster_many(usearch_index_t index, void const* queries, size_t queries_count, size_t queries_stride,
                               usearch_scalar_kind_t query_kind, size_t level, usearch_key_t* keys,
                               usearch_distance_t* distances, uint64_t* computed_distances, uint64_t* visited_members,
                               usearch_error_t* error);

/* `search(exact = true)` of the reference's C++ / Python surface (index.hpp:3047-3051, search_exact_ :4251-4268) for a
 * batch: brute force over every non-removed member of the frozen index, ties resolved exactly like the reference's
 * sequence of sorted inserts (equal distances: larger slot first). Any count; beyond 256 the lists live in L2 instead of registers. */
size_t usearch_b200_exact_search_many(usearch_index_t index, void const* queries, size_t queries_count,
                                      size_t queries_stride, usearch_scalar_kind_t query_kind, size_t count,
                                      usearch_key_t* keys, usearch_distance_t* distances, size_t* counts,
                                      usearch_error_t* error);

/* Phase introspection of the search kernel: enable != 0 turns on (and zeroes) sixteen device-side
 * counters summed over all queries since; `counters16` (may be NULL) first receives the current values:
 * cycles of setup+descent | heap pop | row + visited test | vector wait | distance math | accept replay |
 * output, then queries | heap pushes | sum of per-query max heap size | max heap size | 5 reserved. */
void usearch_b200_profile_phases(usearch_index_t index, int enable, uint64_t* counters16);
/* Tuning knobs of the search launch for this handle ("stage_sets", "warps_per_sm"); results
 * never depend on them. Returns 0, or -1 for an unknown knob. */
int usearch_b200_tune(usearch_index_t index, char const* knob, int value);
int usearch_b200_device(usearch_index_t index);
uint64_t usearch_b200_kernel_launches(usearch_index_t index);
float usearch_b200_last_kernel_ms(usearch_index_t index);
size_t usearch_b200_bytes_per_vector(usearch_index_t index);
size_t usearch_b200_max_level(usearch_index_t index);

#ifdef __cplusplus
}
#endif
#endif /* USEARCH_B200_H */
