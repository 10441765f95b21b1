// Development micro-benchmark (not product code): what read bandwidth does a B200 deliver for RANDOM
// gathers of whole vectors (the HNSW access pattern), as a function of vector size and mechanism?
//   mode 0: warp-per-vector LDG.128 (32 lanes x 16 B per step, U loads in flight per lane)
//   mode 1: 4 lanes per vector LDG.128 (the f32 parity mapping), U loads in flight per lane
//   mode 2: TMA bulk copy (cp.async.bulk) of whole vectors into shared memory, S slots per warp
//   mode 3: sequential streaming read (upper bound for read-only traffic)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bw tools/gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint4 ldg_stream(uint4 const* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int U>
__global__ void k_warp_per_vec(uint8_t const* base, uint32_t nvec, uint32_t vbytes, uint32_t per_warp, uint32_t* sink) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t acc = 0, chunks = vbytes / 16;
    for (uint32_t it = 0; it < per_warp; ++it) {
        uint32_t v = hash32(gw * 7919u + it * 104729u + 12345u) % nvec;
        uint4 const* p = (uint4 const*)(base + (size_t)v * vbytes);
        for (uint32_t j0 = lane; j0 < chunks; j0 += 32 * U) {
            uint4 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 32 < chunks) r[u] = ldg_stream(p + j0 + u * 32);
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 32 < chunks) acc += r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int U>
__global__ void k_quad_per_vec(uint8_t const* base, uint32_t nvec, uint32_t vbytes, uint32_t per_warp, uint32_t* sink) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, g = lane >> 2, sub = lane & 3;
    uint32_t acc = 0, chunks = vbytes / 16;
    for (uint32_t it = 0; it < per_warp; it += 8) {
        uint32_t v = hash32(gw * 7919u + (it + g) * 104729u + 12345u) % nvec;
        uint4 const* p = (uint4 const*)(base + (size_t)v * vbytes);
        for (uint32_t j0 = sub; j0 < chunks; j0 += 4 * U) {
            uint4 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 4 < chunks) r[u] = ldg_stream(p + j0 + u * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 4 < chunks) acc += r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, void const* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one warp per CTA, S slots; each round: issue S bulk copies, wait for all, touch the data lightly
__global__ void k_tma(uint8_t const* base, uint32_t nvec, uint32_t vbytes, uint32_t per_warp, uint32_t slots, uint32_t touch, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t lane = threadIdx.x, gw = blockIdx.x;
    uint32_t stride = (vbytes + 127) / 128 * 128 + 64;
    uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem), bufs = bars + 256;
    if (lane < slots) mbar_init(bars + 8 * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint32_t acc = 0, phase = 0;
    for (uint32_t it = 0; it < per_warp; it += slots) {
        if (lane < slots) {
            uint32_t v = hash32(gw * 7919u + (it + lane) * 104729u + 12345u) % nvec;
            mbar_expect(bars + 8 * lane, vbytes);
            bulk(bufs + lane * stride, base + (size_t)v * vbytes, vbytes, bars + 8 * lane);
        }
        for (uint32_t s = 0; s < slots; ++s) mbar_wait(bars + 8 * s, phase);
        phase ^= 1;
        if (touch) {
            for (uint32_t s = 0; s < slots; ++s) {
                uint4 const* b = (uint4 const*)(smem + 256 + s * stride);
                for (uint32_t j = lane; j < vbytes / 16; j += 32) { uint4 r = b[j]; acc += r.x ^ r.y ^ r.z ^ r.w; }
            }
        }
        __syncwarp();
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// latency probe: same as k_tma but records the average cycles of one round (issue -> all slots landed)
__global__ void k_tma_lat(uint8_t const* base, uint32_t nvec, uint32_t vbytes, uint32_t per_warp, uint32_t slots, uint32_t split,
                          unsigned long long* cyc) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t lane = threadIdx.x, gw = blockIdx.x;
    uint32_t stride = (vbytes + 127) / 128 * 128 + 64;
    uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem), bufs = bars + 256;
    if (lane < slots) mbar_init(bars + 8 * lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;
    unsigned long long total = 0;
    uint32_t part = vbytes / split;
    for (uint32_t it = 0; it < per_warp; it += slots) {
        long long t0 = clock64();
        if (lane < slots * split) {
            uint32_t sl = lane / split, pi = lane % split;
            uint32_t v = hash32(gw * 7919u + (it + sl) * 104729u + 12345u) % nvec;
            if (pi == 0) mbar_expect(bars + 8 * sl, vbytes);
            __syncwarp(__activemask());
            bulk(bufs + sl * stride + pi * part, base + (size_t)v * vbytes + pi * part, part, bars + 8 * sl);
        }
        for (uint32_t s = 0; s < slots; ++s) mbar_wait(bars + 8 * s, phase);
        phase ^= 1;
        total += (unsigned long long)(clock64() - t0);
        __syncwarp();
    }
    if (lane == 0) atomicAdd(cyc, total / (per_warp / slots));
}

template <int U>
__global__ void k_ldg_lat(uint8_t const* base, uint32_t nvec, uint32_t vbytes, uint32_t per_warp, unsigned long long* cyc, uint32_t* sink) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t acc = 0, chunks = vbytes / 16;
    unsigned long long total = 0;
    for (uint32_t it = 0; it < per_warp; ++it) {
        long long t0 = clock64();
        uint32_t v = hash32(gw * 7919u + it * 104729u + 12345u) % nvec;
        uint4 const* p = (uint4 const*)(base + (size_t)v * vbytes);
        for (uint32_t j0 = lane; j0 < chunks; j0 += 32 * U) {
            uint4 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 32 < chunks) r[u] = ldg_stream(p + j0 + u * 32);
#pragma unroll
            for (int u = 0; u < U; ++u) if (j0 + u * 32 < chunks) acc += r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
        }
        if (acc == 0x12345678u) sink[0] = acc;
        total += (unsigned long long)(clock64() - t0);
    }
    if (lane == 0) atomicAdd(cyc, total / per_warp);
}

__global__ void k_stream(uint4 const* base, size_t n16, uint32_t* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = ldg_stream(base + i), b = ldg_stream(base + i + stride), c = ldg_stream(base + i + 2 * stride), d = ldg_stream(base + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main(int argc, char** argv) {
    size_t total = (size_t)3 << 30;
    uint8_t* buf; uint32_t* sink;
    CK(cudaMalloc(&buf, total)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(buf, 1, total));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](char const* name, double bytes, float ms) { printf("%-44s %8.1f GB/s  (%.3f ms)\n", name, bytes / ms / 1e6, ms); fflush(stdout); };
    float ms;
    {   // streaming read
        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_stream<<<sms * 16, 256>>>((uint4 const*)buf, total / 16, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
        cudaEventElapsedTime(&ms, e0, e1); report("stream read 3 GiB", (double)total, ms);
    }
    {   // latency vs load for 3 KB vectors
        unsigned long long* cyc; CK(cudaMalloc(&cyc, 8));
        uint32_t vb = 3072, nvec = (uint32_t)(total / vb);
        for (uint32_t split : {1u, 4u}) for (uint32_t slots : {1u, 8u}) for (int w : {1, 2, 4, 7}) {
            if (slots * split > 32) continue;
            uint32_t stride = (vb + 127) / 128 * 128 + 64; size_t smem = 256 + (size_t)slots * stride;
            CK(cudaFuncSetAttribute(k_tma_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            uint32_t per_warp = 2048 / slots * slots;
            CK(cudaMemset(cyc, 0, 8));
            cudaEventRecord(e0); k_tma_lat<<<sms * w, 32, smem>>>(buf, nvec, vb, per_warp, slots, split, cyc); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("TMA 3072 B split=%u slots=%u warps/SM=%d : %8.1f GB/s, round latency %6.0f cycles\n", split, slots, w,
                   (double)per_warp * sms * w * vb / ms / 1e6, (double)h / (sms * w)); fflush(stdout);
        }
        for (int w : {1, 4, 16}) {
            uint32_t per_warp = 2048;
            CK(cudaMemset(cyc, 0, 8));
            cudaEventRecord(e0); k_ldg_lat<8><<<sms * w / (w >= 4 ? 4 : 1), w >= 4 ? 128 : 32>>>(buf, nvec, vb, per_warp, cyc, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("LDG warp/vec 3072 B warps/SM=%d : %8.1f GB/s, vector latency %6.0f cycles\n", w, (double)per_warp * sms * w * vb / ms / 1e6, (double)h / (sms * w)); fflush(stdout);
        }
    }
    if (argc > 1) return 0;
    uint32_t sizes[] = {3072, 1024, 256, 64};
    for (uint32_t vb : sizes) {
        uint32_t nvec = (uint32_t)(total / vb);
        char name[128];
        for (int warps_per_sm : {16, 32, 48}) {
            uint32_t per_warp = (uint32_t)(((size_t)24 << 30) / vb / (sms * warps_per_sm)); per_warp = per_warp / 8 * 8; if (per_warp > 4096) per_warp = 4096;
            int blocks = sms * warps_per_sm / 4;
            double bytes = (double)per_warp * sms * warps_per_sm * vb;
            for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_warp_per_vec<8><<<blocks, 128>>>(buf, nvec, vb, per_warp, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
            cudaEventElapsedTime(&ms, e0, e1); snprintf(name, 128, "LDG warp/vec U=8  %5u B  %2d warps/SM", vb, warps_per_sm); report(name, bytes, ms);
            for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_quad_per_vec<8><<<blocks, 128>>>(buf, nvec, vb, per_warp, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
            cudaEventElapsedTime(&ms, e0, e1); snprintf(name, 128, "LDG 4-lane/vec U=8 %5u B  %2d warps/SM", vb, warps_per_sm); report(name, bytes, ms);
        }
        if (vb >= 256) {
            for (uint32_t slots : {4u, 8u, 16u}) {
                uint32_t stride = (vb + 127) / 128 * 128 + 64;
                size_t smem = 256 + (size_t)slots * stride;
                CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tma, 32, smem));
                for (int cap : {8, 16, 32}) {
                    int w = per_sm < cap ? per_sm : cap;
                    if (w < cap && cap != 8 && per_sm < cap / 2) continue;
                    uint32_t per_warp = (uint32_t)(((size_t)24 << 30) / vb / (sms * w)); per_warp = per_warp / slots * slots; if (per_warp > 8192) per_warp = 8192;
                    double bytes = (double)per_warp * sms * w * vb;
                    for (uint32_t touch : {0u, 1u}) {
                        for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_tma<<<sms * w, 32, smem>>>(buf, nvec, vb, per_warp, slots, touch, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); }
                        cudaEventElapsedTime(&ms, e0, e1);
                        snprintf(name, 128, "TMA bulk %5u B slots=%2u warps/SM=%2d touch=%u", vb, slots, w, touch); report(name, bytes, ms);
                    }
                }
            }
        }
    }
    return 0;
}
