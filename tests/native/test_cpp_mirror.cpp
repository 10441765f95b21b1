// C++11 client of include/usearch_b200.hpp: the call sites read like code written against the reference's
// index_dense_gt (make / search / dump_to / merge_into / contains), cf. cpp/test.cpp:205-373.
// usage: test_cpp_mirror <index.usearch> <cases.bin>   (same cases.bin as test_c_abi.c)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "usearch_b200.hpp"

using namespace usearch_b200;

#define EXPECT(cond)                                                               \
    do {                                                                           \
        if (!(cond)) {                                                             \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::FILE* f = std::fopen(argv[2], "rb");
    EXPECT(f);
    unsigned long long nq, dims, k;
    EXPECT(std::fread(&nq, 8, 1, f) == 1 && std::fread(&dims, 8, 1, f) == 1 && std::fread(&k, 8, 1, f) == 1);
    std::vector<float> queries(nq * dims), want_dist(nq * k);
    std::vector<vector_key_t> want_keys(nq * k);
    std::vector<unsigned long long> want_counts(nq);
    EXPECT(std::fread(queries.data(), 4, nq * dims, f) == nq * dims);
    EXPECT(std::fread(want_keys.data(), 8, nq * k, f) == nq * k);
    EXPECT(std::fread(want_dist.data(), 4, nq * k, f) == nq * k);
    EXPECT(std::fread(want_counts.data(), 8, nq, f) == nq);
    std::fclose(f);

    index_dense_t::state_result_t state = index_dense_t::make(argv[1]);
    EXPECT(state);
    index_dense_t& index = state.index;
    EXPECT(index.size() > 0 && index.dimensions() == dims);

    index_dense_t::search_result_t result = index.search(queries.data(), k);
    EXPECT(result && result.size() == want_counts[0]);
    EXPECT(result.computed_distances > 0 && result.visited_members > 0);
    for (std::size_t i = 0; i != result.size(); ++i) {
        EXPECT(result[i].member.key == want_keys[i]);
        distance_t const got = result[i].distance;
        EXPECT(std::memcmp(&got, &want_dist[i], 4) == 0);
        EXPECT(result.contains(want_keys[i]));
    }
    EXPECT(result.front().distance <= result.back().distance);

    // dump_to pads like index.hpp:2715-2720
    std::vector<vector_key_t> keys(k + 2, 99);
    std::vector<distance_t> dist(k + 2, 0.f);
    EXPECT(result.dump_to(keys.data(), dist.data(), k + 2) == result.size());
    EXPECT(keys[k + 1] == 0 && std::isnan(dist[k + 1]));

    // merge_into: merging a result into itself keeps the best k with ties (index.hpp:2650-2670)
    std::vector<vector_key_t> merged_keys(k);
    std::vector<distance_t> merged_dist(k);
    std::size_t merged = result.merge_into(merged_keys.data(), merged_dist.data(), 0, k);
    EXPECT(merged == result.size());
    for (std::size_t i = 1; i < merged; ++i) EXPECT(merged_dist[i - 1] <= merged_dist[i]);

    index_dense_t::batch_result_t batch = index.search_many(queries.data(), nq, k);
    EXPECT(batch);
    for (unsigned long long q = 0; q != nq; ++q) {
        EXPECT(batch.counts[q] == want_counts[q]);
        EXPECT(std::memcmp(&batch.keys[q * k], &want_keys[q * k], k * 8) == 0);
        EXPECT(std::memcmp(&batch.distances[q * k], &want_dist[q * k], k * 4) == 0);
    }

    // exact=true scans every member: never worse than the graph search, same first hit on an easy query
    index_dense_t::search_result_t scanned = index.search(queries.data(), k, 0, true);
    EXPECT(scanned && scanned.size() == result.size());
    EXPECT(scanned[0].distance <= result[0].distance);
    for (std::size_t i = 1; i < scanned.size(); ++i) EXPECT(scanned[i - 1].distance <= scanned[i].distance);

    // cluster(vector, level): level beyond the top of the graph -> the entry point, at any query
    index_dense_t::cluster_result_t top = index.cluster(queries.data(), index.max_level() + 1);
    index_dense_t::cluster_result_t top2 = index.cluster(queries.data() + dims, index.max_level() + 1);
    EXPECT(top && top2 && top.cluster.member.key == top2.cluster.member.key);
    EXPECT(top.computed_distances == 2 && top.visited_members == 0);
    index_dense_t::cluster_result_t base = index.cluster(queries.data(), 0);
    EXPECT(base && base.cluster.distance <= top.cluster.distance);

    // make() from a metric, like index_dense_gt::make(metric_punned_t, config) — unsupported pairs fail by value
    index_dense_t::state_result_t fresh = index_dense_t::make(metric_punned_t::builtin(dims, usearch_metric_cos_k, usearch_scalar_f32_k));
    EXPECT(fresh && fresh.index.size() == 0);
    index_dense_t::state_result_t bad = index_dense_t::make(metric_punned_t::builtin(dims, usearch_metric_haversine_k, usearch_scalar_f32_k));
    EXPECT(!bad);

    // the mutation half of index_dense_gt on an index made from a metric: reserve / add / add_many / contains / count / get /
    // rename / remove, then the new members are found by the search (cpp/test.cpp:358-361: a member is its own nearest)
    {
        index_dense_t& built = fresh.index;
        EXPECT(built.try_reserve(nq + 4));
        EXPECT(built.add(1000, queries.data()) && built.size() == 1);
        std::vector<vector_key_t> keys2;
        for (unsigned long long q = 1; q != nq; ++q) keys2.push_back(1000 + q);
        index_dense_t::add_result_t added = built.add_many(keys2.data(), queries.data() + dims, nq - 1);
        EXPECT(added && added.new_size == nq && built.size() == nq);
        EXPECT(!built.add(1000, queries.data()));                       // duplicate key on a non-multi index
        EXPECT(built.contains(1003) && built.count(1003) == 1 && !built.contains(7));
        std::vector<float> stored(dims);
        EXPECT(built.get(1003, stored.data()) == 1 && std::memcmp(stored.data(), queries.data() + 3 * dims, dims * 4) == 0);
        index_dense_t::search_result_t self = built.search(queries.data() + 5 * dims, 3);
        EXPECT(self && self.size() >= 1 && self[0].member.key == 1005 && self[0].distance < 1e-5f);
        EXPECT(built.rename(1005, 2005).completed == 1 && built.contains(2005) && !built.contains(1005));
        EXPECT(built.remove(2005).completed == 1 && built.size() == nq - 1);
        self = built.search(queries.data() + 5 * dims, 3);
        for (std::size_t i = 0; i != self.size(); ++i) EXPECT(self[i].member.key != 2005);
    }

    std::printf("CPP_MIRROR_OK %llu queries\n", nq);
    return 0;
}
