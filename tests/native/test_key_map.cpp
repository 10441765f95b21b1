// Native unit test of the host-side key -> slot table (usearch_b200/csrc/key_map.h): no CUDA, no GPU.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

#include "key_map.h"

using namespace usearch_b200;

#define EXPECT(cond)                                                                 \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main() {
    uint64_t const free_key = ~0ull;
    std::vector<uint64_t> keys;
    key_map_t map;
    EXPECT(!map.contains(5) && map.count(5) == 0); // never built: empty

    // a multi-index: every key three times, some entries removed (free key) before the table is built
    std::mt19937_64 rng(7);
    std::multimap<uint64_t, uint32_t> model;
    for (uint32_t s = 0; s < 30000; ++s) {
        uint64_t key = (s % 10000) * 0x9E3779B97F4A7C15ull; // clustered hashes stress the probing
        if (s % 17 == 0) key = free_key;
        keys.push_back(key);
        if (key != free_key) model.emplace(key, s);
    }
    map.rebuild(keys, free_key, 8); // a deliberately wrong size hint: the table must grow on its own
    EXPECT(map.built);
    for (uint64_t probe = 0; probe < 10000; ++probe) {
        uint64_t const key = probe * 0x9E3779B97F4A7C15ull;
        EXPECT(map.count(key) == model.count(key));
        EXPECT(map.contains(key) == (model.count(key) != 0));
        std::vector<uint32_t> slots;
        map.for_each(key, [&](uint32_t slot, size_t) { slots.push_back(slot); return true; });
        for (uint32_t slot : slots) EXPECT(keys[slot] == key);
    }
    EXPECT(!map.contains(free_key) && !map.contains(12345));

    // incremental inserts (growth re-inserts the live cells), erase through the cell index, re-insert under a new key
    for (uint32_t s = 30000; s < 200000; ++s) {
        uint64_t const key = rng() | 1;
        keys.push_back(key);
        map.insert(key, s);
        model.emplace(key, s);
    }
    for (auto it = model.begin(); it != model.end();) {
        if (rng() % 3 == 0) {
            uint64_t const key = it->first;
            map.for_each(key, [&](uint32_t slot, size_t cell) { keys[slot] = free_key, map.erase_cell(cell); return true; });
            it = model.erase(model.lower_bound(key), model.upper_bound(key));
        } else
            ++it;
    }
    size_t live = 0;
    for (auto const& kv : model) {
        EXPECT(map.contains(kv.first));
        EXPECT(map.count(kv.first) == model.count(kv.first));
        ++live;
    }
    for (uint32_t s = 0; s < keys.size(); ++s)
        if (keys[s] == free_key) EXPECT(true);
    // early exit of for_each
    uint64_t const triple = 3 * 0x9E3779B97F4A7C15ull;
    if (model.count(triple) > 1) {
        size_t seen = 0;
        map.for_each(triple, [&](uint32_t, size_t) { ++seen; return false; });
        EXPECT(seen == 1);
    }
    map.clear();
    EXPECT(!map.built && !map.contains(triple));
    std::printf("KEY_MAP_OK %zu live entries\n", live);
    return 0;
}
