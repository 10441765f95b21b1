/*
 *  oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
 *
 *  A thin driver around the UNMODIFIED reference headers. It is compiled against
 *  /root/reference/include (never copied into this repository) by oracle/build.py into
 *  oracle/_ref/libusearch_ref.so and gives the parity tests, `__graft_entry__.smoke()` and
 *  bench.py's CPU arms three things the reference's C ABI (c/usearch.h) does not expose:
 *
 *    1. a thread-pool batch search that also returns the per-query `computed_distances` and
 *       `visited_members` counters (index.hpp:2605-2609) — those counters define the
 *       algorithmic bytes of the roofline;
 *    2. the ability to pin the metric to the portable restatement in metrics_pinned.h through
 *       `metric_punned_t::stateless` (index_plugins.hpp:1772-1786), so that the traversal of
 *       `index_gt::search` (index.hpp:3016-3075) runs on arithmetic a GPU can reproduce
 *       bit-for-bit, independent of which SimSIMD kernel the host CPU would select;
 *    3. multi-threaded index construction + serialisation to the v2 buffer format, which is the
 *       only hand-off between reference and device ("same serialized graph" parity policy).
 *
 *  Everything that decides *which labels come back* — descent, best-first expansion, containers,
 *  tie-breaking — is the reference's own code. Only the distance function is swappable.
 */
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <usearch/index_dense.hpp>

#include "metrics_pinned.h"

using namespace unum::usearch;

namespace {

using index_t = index_dense_gt<std::uint64_t, std::uint32_t>;

struct ref_index_t {
    index_t index;
    std::size_t threads = 0;
};

template <typename fn_at> std::uintptr_t fn_addr(fn_at fn) { return reinterpret_cast<std::uintptr_t>(fn); }

/* metric_punned_t::invoke_array_array_third passes (a, b, third) as three uptr_t values */
#define PIN3(name, type)                                                                                     \
    static float pin_##name(std::size_t a, std::size_t b, std::size_t n) {                                   \
        return pinned_##name(reinterpret_cast<type const*>(a), reinterpret_cast<type const*>(b), n);         \
    }
PIN3(l2sq_f32, float)
PIN3(ip_f32, float)
PIN3(cos_f32, float)
PIN3(l2sq_f16, std::uint16_t)
PIN3(ip_f16, std::uint16_t)
PIN3(cos_f16, std::uint16_t)
PIN3(l2sq_bf16, std::uint16_t)
PIN3(ip_bf16, std::uint16_t)
PIN3(cos_bf16, std::uint16_t)
PIN3(l2sq_i8, std::int8_t)
PIN3(ip_i8, std::int8_t)
PIN3(cos_i8, std::int8_t)
PIN3(hamming_b1, std::uint8_t)
PIN3(tanimoto_b1, std::uint8_t)
PIN3(sorensen_b1, std::uint8_t)

std::uintptr_t pinned_for(metric_kind_t m, scalar_kind_t s) {
    switch (s) {
    case scalar_kind_t::f32_k:
        if (m == metric_kind_t::l2sq_k) return fn_addr(&pin_l2sq_f32);
        if (m == metric_kind_t::ip_k) return fn_addr(&pin_ip_f32);
        if (m == metric_kind_t::cos_k) return fn_addr(&pin_cos_f32);
        break;
    case scalar_kind_t::f16_k:
        if (m == metric_kind_t::l2sq_k) return fn_addr(&pin_l2sq_f16);
        if (m == metric_kind_t::ip_k) return fn_addr(&pin_ip_f16);
        if (m == metric_kind_t::cos_k) return fn_addr(&pin_cos_f16);
        break;
    case scalar_kind_t::bf16_k:
        if (m == metric_kind_t::l2sq_k) return fn_addr(&pin_l2sq_bf16);
        if (m == metric_kind_t::ip_k) return fn_addr(&pin_ip_bf16);
        if (m == metric_kind_t::cos_k) return fn_addr(&pin_cos_bf16);
        break;
    case scalar_kind_t::i8_k:
        if (m == metric_kind_t::l2sq_k) return fn_addr(&pin_l2sq_i8);
        if (m == metric_kind_t::ip_k) return fn_addr(&pin_ip_i8);
        if (m == metric_kind_t::cos_k) return fn_addr(&pin_cos_i8);
        break;
    case scalar_kind_t::b1x8_k:
        if (m == metric_kind_t::hamming_k) return fn_addr(&pin_hamming_b1);
        if (m == metric_kind_t::tanimoto_k || m == metric_kind_t::jaccard_k) return fn_addr(&pin_tanimoto_b1);
        if (m == metric_kind_t::sorensen_k) return fn_addr(&pin_sorensen_b1);
        break;
    default: break;
    }
    return 0;
}

template <typename fn_at> void parallel_for(std::size_t n, std::size_t threads, fn_at&& fn) {
    if (threads <= 1 || n <= 1) {
        for (std::size_t i = 0; i < n; ++i) fn(0, i);
        return;
    }
    std::atomic<std::size_t> cursor{0};
    std::vector<std::thread> pool;
    std::size_t const grain = 16;
    for (std::size_t t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (;;) {
                std::size_t begin = cursor.fetch_add(grain);
                if (begin >= n) break;
                std::size_t end = begin + grain < n ? begin + grain : n;
                for (std::size_t i = begin; i < end; ++i) fn(t, i);
            }
        });
    for (auto& th : pool) th.join();
}

bool ensure_threads(ref_index_t* r, std::size_t members, std::size_t threads) {
    if (threads == 0) threads = 1;
    if (r->threads >= threads && r->index.capacity() >= members) return true;
    std::size_t cap = r->index.capacity() > members ? r->index.capacity() : members;
    index_limits_t limits(cap, threads > r->threads ? threads : r->threads);
    if (!r->index.try_reserve(limits)) return false;
    r->threads = limits.threads();
    return true;
}

} // namespace

extern "C" {

void* ref_make(int metric_char, int scalar_char, std::size_t dimensions, std::size_t connectivity,
               std::size_t expansion_add, std::size_t expansion_search, char const** error) {
    *error = nullptr;
    index_dense_config_t config;
    config.connectivity = connectivity;
    config.connectivity_base = connectivity * 2; /* index.hpp:1368 default, made explicit */
    config.expansion_add = expansion_add;
    config.expansion_search = expansion_search;
    config.enable_key_lookups = true;
    metric_punned_t metric = metric_punned_t::builtin(dimensions, static_cast<metric_kind_t>(metric_char),
                                                      static_cast<scalar_kind_t>(scalar_char));
    if (metric.missing()) {
        *error = "Unknown metric kind!";
        return nullptr;
    }
    auto state = index_t::make(metric, config);
    if (!state) {
        *error = state.error.release();
        return nullptr;
    }
    auto* r = new ref_index_t{std::move(state.index), 0};
    return r;
}

void* ref_make_empty(void) { return new ref_index_t{}; }

void ref_free(void* h) { delete static_cast<ref_index_t*>(h); }

std::size_t ref_size(void* h) { return static_cast<ref_index_t*>(h)->index.size(); }
std::size_t ref_dimensions(void* h) { return static_cast<ref_index_t*>(h)->index.dimensions(); }
std::size_t ref_connectivity(void* h) { return static_cast<ref_index_t*>(h)->index.connectivity(); }
std::size_t ref_max_level(void* h) { return static_cast<ref_index_t*>(h)->index.max_level(); }
std::size_t ref_bytes_per_vector(void* h) { return static_cast<ref_index_t*>(h)->index.bytes_per_vector(); }
std::size_t ref_expansion_search(void* h) { return static_cast<ref_index_t*>(h)->index.expansion_search(); }
void ref_change_expansion_search(void* h, std::size_t ef) {
    static_cast<ref_index_t*>(h)->index.change_expansion_search(ef);
}
int ref_metric_kind(void* h) { return static_cast<int>(static_cast<ref_index_t*>(h)->index.metric().metric_kind()); }
int ref_scalar_kind(void* h) { return static_cast<int>(static_cast<ref_index_t*>(h)->index.metric().scalar_kind()); }
char const* ref_isa_name(void* h) { return static_cast<ref_index_t*>(h)->index.metric().isa_name(); }

/* mode 0: the reference's own builtin metric (native SimSIMD dispatch on this host);
 * mode 1: the portable pinned restatement (metrics_pinned.h). Returns 0 on success. */
int ref_pin_metric(void* h, int mode) {
    auto* r = static_cast<ref_index_t*>(h);
    metric_punned_t const& old = r->index.metric();
    metric_kind_t m = old.metric_kind();
    scalar_kind_t s = old.scalar_kind();
    std::size_t d = old.dimensions();
    if (mode == 0) {
        r->index.change_metric(metric_punned_t::builtin(d, m, s));
        return 0;
    }
#if defined(__FAST_MATH__)
    return -2; /* the pinned arithmetic is only exact without -ffast-math: use the parity flavour */
#endif
    std::uintptr_t fn = pinned_for(m, s);
    if (!fn) return -1;
    r->index.change_metric(metric_punned_t::stateless(d, fn, metric_punned_signature_t::array_array_size_k, m, s));
    return 0;
}

/* index_dense_gt::cluster(vector, level) (index_dense.hpp:788-793) for a batch, single-threaded; queries in the index's
 * scalar kind. */
void ref_cluster_many(void* h, void const* queries, std::size_t nq, std::size_t stride_bytes, std::size_t level,
                      std::uint64_t* keys, float* distances, std::uint64_t* computed, std::uint64_t* visited, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    if (!ensure_threads(r, r->index.size(), 1)) {
        *error = "Out of memory!";
        return;
    }
    for (std::size_t i = 0; i != nq; ++i) {
        byte_t const* q = static_cast<byte_t const*>(queries) + i * stride_bytes;
        index_t::cluster_result_t result;
        switch (r->index.metric().scalar_kind()) {
        case scalar_kind_t::f32_k: result = r->index.cluster(reinterpret_cast<f32_t const*>(q), level, 0); break;
        case scalar_kind_t::f16_k: result = r->index.cluster(reinterpret_cast<f16_t const*>(q), level, 0); break;
        case scalar_kind_t::bf16_k: result = r->index.cluster(reinterpret_cast<bf16_t const*>(q), level, 0); break;
        case scalar_kind_t::i8_k: result = r->index.cluster(reinterpret_cast<i8_t const*>(q), level, 0); break;
        case scalar_kind_t::b1x8_k: result = r->index.cluster(reinterpret_cast<b1x8_t const*>(q), level, 0); break;
        default: *error = "Unsupported scalar kind"; return;
        }
        if (!result) {
            *error = "cluster failed";
            return;
        }
        keys[i] = result.cluster.member.key;
        distances[i] = result.cluster.distance;
        computed[i] = result.computed_distances;
        visited[i] = result.visited_members;
    }
}

/* exact_search_t (index_plugins.hpp:2071-2164) as usearch_exact_search (c/lib.cpp:468-501) drives it, single-threaded.
 * pinned != 0 swaps the metric for the portable restatement. Outputs are dense [nq x wanted]. Returns 0 on success. */
int ref_exact_search(void const* dataset, std::size_t n, std::size_t dataset_stride, void const* queries, std::size_t nq,
                     std::size_t queries_stride, int metric_char, int scalar_char, std::size_t dimensions, std::size_t wanted,
                     int pinned, std::uint64_t* keys, float* distances) {
    metric_kind_t m = static_cast<metric_kind_t>(metric_char);
    scalar_kind_t s = static_cast<scalar_kind_t>(scalar_char);
    metric_punned_t metric = metric_punned_t::builtin(dimensions, m, s);
    if (metric.missing()) return -1;
    if (pinned) {
#if defined(__FAST_MATH__)
        return -2;
#endif
        std::uintptr_t fn = pinned_for(m, s);
        if (!fn) return -1;
        metric = metric_punned_t::stateless(dimensions, fn, metric_punned_signature_t::array_array_size_k, m, s);
    }
    exact_search_t search;
    exact_search_results_t result = search(static_cast<byte_t const*>(dataset), n, dataset_stride,
                                           static_cast<byte_t const*>(queries), nq, queries_stride, wanted, metric);
    if (!result) return -3;
    for (std::size_t q = 0; q != nq; ++q) {
        auto row = result.at(q);
        for (std::size_t i = 0; i != wanted; ++i)
            keys[q * wanted + i] = row[i].offset, distances[q * wanted + i] = row[i].distance;
    }
    return 0;
}

/* One distance through whatever metric is currently installed (a, b in the index's scalar kind). */
float ref_distance(void* h, void const* a, void const* b) {
    auto* r = static_cast<ref_index_t*>(h);
    return r->index.metric()(static_cast<byte_t const*>(a), static_cast<byte_t const*>(b));
}

/* Vectors are given in the index's own scalar kind (no cast on the way in). */
std::size_t ref_add_many(void* h, std::uint64_t const* keys, void const* vectors, std::size_t n,
                         std::size_t stride_bytes, std::size_t threads, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    if (!ensure_threads(r, r->index.size() + n, threads)) {
        *error = "Out of memory!";
        return 0;
    }
    scalar_kind_t s = r->index.metric().scalar_kind();
    std::atomic<std::size_t> done{0};
    std::atomic<char const*> first_error{nullptr};
    auto const* base = static_cast<byte_t const*>(vectors);
    parallel_for(n, threads, [&](std::size_t thread, std::size_t i) {
        byte_t const* v = base + i * stride_bytes;
        index_t::add_result_t result;
        switch (s) {
        case scalar_kind_t::f32_k: result = r->index.add(keys[i], reinterpret_cast<f32_t const*>(v), thread); break;
        case scalar_kind_t::f16_k: result = r->index.add(keys[i], reinterpret_cast<f16_t const*>(v), thread); break;
        case scalar_kind_t::bf16_k: result = r->index.add(keys[i], reinterpret_cast<bf16_t const*>(v), thread); break;
        case scalar_kind_t::i8_k: result = r->index.add(keys[i], reinterpret_cast<i8_t const*>(v), thread); break;
        case scalar_kind_t::b1x8_k: result = r->index.add(keys[i], reinterpret_cast<b1x8_t const*>(v), thread); break;
        default: result = index_t::add_result_t{}.failed("Unsupported scalar kind!"); break;
        }
        if (!result) {
            char const* expected = nullptr;
            first_error.compare_exchange_strong(expected, result.error.release());
        } else
            done.fetch_add(1);
    });
    *error = first_error.load();
    return done.load();
}

std::size_t ref_remove(void* h, std::uint64_t key, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    auto result = r->index.remove(key);
    if (!result) {
        *error = result.error.release();
        return 0;
    }
    return result.completed;
}

std::size_t ref_serialized_length(void* h) { return static_cast<ref_index_t*>(h)->index.serialized_length(); }

void ref_save_buffer(void* h, void* buffer, std::size_t length, char const** error) {
    *error = nullptr;
    memory_mapped_file_t map(static_cast<byte_t*>(buffer), length);
    auto result = static_cast<ref_index_t*>(h)->index.save(std::move(map));
    if (!result) *error = result.error.release();
}

void ref_load_buffer(void* h, void const* buffer, std::size_t length, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    memory_mapped_file_t map(static_cast<byte_t*>(const_cast<void*>(buffer)), length);
    auto result = r->index.load(std::move(map));
    if (!result) {
        *error = result.error.release();
        return;
    }
    r->threads = 0;
}

/* index_dense_gt::view over caller memory (index_dense.hpp:1190-1313): no copy of the vectors; the caller keeps `buffer` alive */
void ref_view_buffer(void* h, void const* buffer, std::size_t length, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    memory_mapped_file_t map(static_cast<byte_t*>(const_cast<void*>(buffer)), length);
    auto result = r->index.view(std::move(map));
    if (!result) {
        *error = result.error.release();
        return;
    }
    r->threads = 0;
}

void ref_save_path(void* h, char const* path, char const** error) {
    *error = nullptr;
    auto result = static_cast<ref_index_t*>(h)->index.save(path);
    if (!result) *error = result.error.release();
}

void ref_load_path(void* h, char const* path, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    auto result = r->index.load(path);
    if (!result) {
        *error = result.error.release();
        return;
    }
    r->threads = 0;
}

/*
 *  The reference's batch search: a pool of threads, each running independent single-query
 *  `index_dense_gt::search` calls (python/lib.cpp:286-308 does exactly this). Queries are in the
 *  index's scalar kind. Outputs follow `dump_to` (index.hpp:2707-2722): unused slots hold key 0
 *  and a NaN distance. `computed`/`visited` may be NULL.
 */
void ref_search_many(void* h, void const* queries, std::size_t nq, std::size_t stride_bytes, std::size_t wanted,
                     std::size_t threads, int exact, std::uint64_t* keys, float* distances, std::uint64_t* counts,
                     std::uint64_t* computed, std::uint64_t* visited, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    if (!ensure_threads(r, r->index.size(), threads)) {
        *error = "Out of memory!";
        return;
    }
    scalar_kind_t s = r->index.metric().scalar_kind();
    std::atomic<char const*> first_error{nullptr};
    auto const* base = static_cast<byte_t const*>(queries);
    parallel_for(nq, threads, [&](std::size_t thread, std::size_t i) {
        byte_t const* q = base + i * stride_bytes;
        bool ok = true;
        auto consume = [&](index_t::search_result_t&& result) {
            if (!result) {
                char const* expected = nullptr;
                first_error.compare_exchange_strong(expected, result.error.release());
                ok = false;
                counts[i] = 0;
                return;
            }
            counts[i] = result.dump_to(keys + i * wanted, distances + i * wanted, wanted);
            if (computed) computed[i] = result.computed_distances;
            if (visited) visited[i] = result.visited_members;
        };
        switch (s) {
        case scalar_kind_t::f32_k: consume(r->index.search(reinterpret_cast<f32_t const*>(q), wanted, thread, exact)); break;
        case scalar_kind_t::f16_k: consume(r->index.search(reinterpret_cast<f16_t const*>(q), wanted, thread, exact)); break;
        case scalar_kind_t::bf16_k: consume(r->index.search(reinterpret_cast<bf16_t const*>(q), wanted, thread, exact)); break;
        case scalar_kind_t::i8_k: consume(r->index.search(reinterpret_cast<i8_t const*>(q), wanted, thread, exact)); break;
        case scalar_kind_t::b1x8_k: consume(r->index.search(reinterpret_cast<b1x8_t const*>(q), wanted, thread, exact)); break;
        default: {
            char const* expected = nullptr;
            first_error.compare_exchange_strong(expected, "Unsupported scalar kind!");
        }
        }
        (void)ok;
    });
    *error = first_error.load();
}

/* `filtered_search` (index_dense.hpp:774-779) with the predicate "key is in the sorted array `allowed`";
 * f32-only convenience for the parity tests of the device-side filter. */
void ref_filtered_search_many_f32(void* h, float const* queries, std::size_t nq, std::size_t stride_bytes,
                                  std::size_t wanted, std::size_t threads, std::uint64_t const* allowed,
                                  std::size_t allowed_count, std::uint64_t* keys, float* distances, std::uint64_t* counts,
                                  std::uint64_t* computed, std::uint64_t* visited, char const** error) {
    *error = nullptr;
    auto* r = static_cast<ref_index_t*>(h);
    if (!ensure_threads(r, r->index.size(), threads)) {
        *error = "Out of memory!";
        return;
    }
    auto const* base = reinterpret_cast<byte_t const*>(queries);
    auto predicate = [=](std::uint64_t key) noexcept { return std::binary_search(allowed, allowed + allowed_count, key); };
    parallel_for(nq, threads, [&](std::size_t thread, std::size_t i) {
        auto result = r->index.filtered_search(reinterpret_cast<f32_t const*>(base + i * stride_bytes), wanted, predicate, thread);
        counts[i] = result.dump_to(keys + i * wanted, distances + i * wanted, wanted);
        if (computed) computed[i] = result.computed_distances;
        if (visited) visited[i] = result.visited_members;
    });
}

/* The reference's query-side casts (index_plugins.hpp:1105-1224), exposed so tests can check the
 * device-side casts against them: f32 → the index's scalar kind. Returns bytes written. */
std::size_t ref_cast_from_f32(int scalar_char, float const* input, std::size_t dimensions, void* output) {
    scalar_kind_t s = static_cast<scalar_kind_t>(scalar_char);
    casts_punned_t casts = casts_punned_t::make(s);
    std::size_t bytes = (dimensions * bits_per_scalar(s) + 7) / 8;
    bool casted = casts.from.f32(reinterpret_cast<byte_t const*>(input), dimensions, static_cast<byte_t*>(output));
    if (!casted) std::memcpy(output, input, bytes);
    return bytes;
}

std::size_t ref_hardware_threads(void) { return std::thread::hardware_concurrency(); }

} // extern "C"
