// Write-only HBM bandwidth on B200 for different store flavours (the feature kernel is a pure write stream).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/hbm_write_variants profiles/hbm_write_variants.cu
//   ./profiles/hbm_write_variants            (prints GB/s per variant, 4 GiB buffer, median of 10)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

enum { ST_DEFAULT = 0, ST_CS = 1, ST_WT = 2, ST_NOALLOC = 3, ST_EVICT_FIRST = 4 };

template <int MODE>
__device__ __forceinline__ void store16(uint4* p, uint4 v) {
    if (MODE == ST_DEFAULT) *p = v;
    else if (MODE == ST_CS) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == ST_WT) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == ST_NOALLOC) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else {      // 32-byte stores only: sm_100 accepts the L2 eviction-priority qualifier on .v8.b32 / .v4.b64
        asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// 32 B per thread (st.global.v8.b32), default policy or evict_first
template <int EF>
__global__ void sweep32_kernel(uint4* out, size_t n32) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
        uint4* p = out + 2 * i;
        if (EF) asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// grid-stride: consecutive threads write consecutive 16 B (fully coalesced), the whole grid sweeps the buffer
template <int MODE>
__global__ void sweep_kernel(uint4* out, size_t n16) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        store16<MODE>(out + i, v);
}

// every CTA owns contiguous chunks of `chunk16` x 16 B (like one window's feature block), CTAs walk chunk by chunk
template <int MODE>
__global__ void chunk_kernel(uint4* out, size_t n16, size_t chunk16) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
    const size_t n_chunks = n16 / chunk16;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint4* p = out + c * chunk16;
        for (size_t i = threadIdx.x; i < chunk16; i += blockDim.x) store16<MODE>(p + i, v);
    }
}

// TMA bulk store: stage 16 KB in shared memory once, then cp.async.bulk shared -> global per chunk
__global__ void bulk_kernel(uint8_t* out, size_t nbytes, int chunk) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t n_chunks = nbytes / chunk;
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        int inflight = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + c * (size_t)chunk), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    std::vector<float> ts;
    for (int i = 0; i < 10; ++i) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main() {
    const size_t N = (size_t)4 << 30;
    uint8_t* buf; CK(cudaMalloc(&buf, N));
    uint4* b16 = reinterpret_cast<uint4*>(buf);
    const size_t n16 = N / 16;
    const char* names[] = {"default", "st.cs", "st.wt", "L1::no_allocate", "L2::evict_first"};
    for (int grid : {148 * 2, 148 * 8, 148 * 32}) {
        float ms[4];
        ms[0] = time_it([&] { sweep_kernel<0><<<grid, 256>>>(b16, n16); });
        ms[1] = time_it([&] { sweep_kernel<1><<<grid, 256>>>(b16, n16); });
        ms[2] = time_it([&] { sweep_kernel<2><<<grid, 256>>>(b16, n16); });
        ms[3] = time_it([&] { sweep_kernel<3><<<grid, 256>>>(b16, n16); });
        for (int k = 0; k < 4; ++k) printf("sweep  grid %5d  %-16s %8.1f GB/s\n", grid, names[k], N / ms[k] / 1e6);
        float a = time_it([&] { sweep32_kernel<0><<<grid, 256>>>(b16, n16 / 2); });
        float b = time_it([&] { sweep32_kernel<1><<<grid, 256>>>(b16, n16 / 2); });
        printf("sweep  grid %5d  32 B stores      %8.1f GB/s   32 B + L2::evict_first %8.1f GB/s\n", grid, N / a / 1e6, N / b / 1e6);
    }
    for (size_t chunk : {(size_t)16384, (size_t)147456, (size_t)294912}) {      // 16 KB, one S=48 plane pair row group, one window
        const size_t c16 = chunk / 16;
        float m0 = time_it([&] { chunk_kernel<0><<<148 * 2, 256>>>(b16, n16, c16); });
        float m1 = time_it([&] { chunk_kernel<1><<<148 * 2, 256>>>(b16, n16, c16); });
        printf("chunk %7zu B  grid 296  default %8.1f  st.cs %8.1f GB/s\n", chunk, N / m0 / 1e6, N / m1 / 1e6);
        float m2 = time_it([&] { chunk_kernel<0><<<148 * 8, 256>>>(b16, n16, c16); });
        printf("chunk %7zu B  grid 1184 default %8.1f GB/s\n", chunk, N / m2 / 1e6);
    }
    for (int chunk : {4096, 16384, 65536}) {
        for (int grid : {148, 148 * 2, 148 * 4}) {
            CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk));
            float ms = time_it([&] { bulk_kernel<<<grid, 128, chunk>>>(buf, N, chunk); });
            printf("TMA bulk store  chunk %6d B  grid %4d  %8.1f GB/s\n", chunk, grid, N / ms / 1e6);
        }
    }
    CK(cudaMemset(buf, 0, N));
    float msm = time_it([&] { cudaMemsetAsync(buf, 1, N); });
    printf("cudaMemsetAsync %8.1f GB/s\n", N / msm / 1e6);
    CK(cudaDeviceSynchronize());
    return 0;
}
