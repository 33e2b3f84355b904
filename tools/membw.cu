// membw.cu — development probe: what does THIS B200 sustain for write-only / read-only / copy streams?
// (The roofline denominator stays MEASURED_PEAKS.json's copy figure; this only tells how far a write-dominated
//  kernel such as the fused parse pass can possibly get.)   nvcc -arch=sm_100a -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k_write(uint4* p, size_t n, int cs) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (cs) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else p[i] = v;
    }
}
__global__ void k_read(const uint4* p, size_t n, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
// write through shared memory + TMA bulk stores: each CTA streams `chunk`-byte blocks
__global__ void k_bulk_write(uint8_t* p, size_t nbytes, uint32_t chunk) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (uint32_t i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        for (size_t o = (size_t)blockIdx.x * chunk; o + chunk <= nbytes; o += (size_t)gridDim.x * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + o), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
// emulate the fused pass's label tiles: CTAs draw regions of `per_region` chunks from a counter; the 8 warps of a CTA
// push the region's chunks round-robin (warp w: chunks w, w+8, ...) with `depth` bulk stores in flight per warp
__global__ void k_bulk_pattern(uint8_t* p, size_t nbytes, uint32_t chunk, uint32_t per_region, uint32_t misalign,
                               unsigned int* counter, int depth) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ unsigned int s_region;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < 8 * 2 * chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t region_bytes = (size_t)chunk * per_region;
    const size_t n_regions = (nbytes - 256) / region_bytes;
    uint32_t it = 0;
    for (;;) {
        if (threadIdx.x == 0) s_region = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int region = s_region;
        if (region >= n_regions) break;
        uint8_t* base = p + region * region_bytes + misalign;
        for (uint32_t c = warp; c < per_region; c += 8, it++) {
            if (depth >= 10) {   // as the fused pass: every lane writes the block, then a proxy fence, then the store
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float* blk = reinterpret_cast<float*>(sm + (warp * 2 + (it & 1)) * chunk);
                blk[(lane * 20 + it) % (chunk / 4)] = 1.0f;
                blk[(lane * 20 + 10 + it) % (chunk / 4)] = 0.0f;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            }
            if (lane == 0) {
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + (warp * 2 + (it & 1)) * chunk);
                if (depth >= 10) {}
                else if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)c * chunk), "r"(s), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// the fused parse pass's traffic without its arithmetic: per "record" 32 tiles of 8 KiB are TMA-loaded; 24 of them
// ("image") are answered with 32 KiB of streaming float4 stores, 8 ("label") with 320 KiB of 2560-byte bulk stores
__global__ void k_mix(const uint8_t* in, uint8_t* outA, uint8_t* outB, uint32_t n_records, unsigned int* counter) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned int s_chunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = sm;
    float* hot = reinterpret_cast<float*>(sm + 8320) + warp * 2 * 640;
    for (uint32_t i = threadIdx.x; i < 8 * 2 * 640; i += blockDim.x) reinterpret_cast<float*>(sm + 8320)[i] = 0.f;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0, hit = 0;
    const uint32_t total = n_records * 32;
    for (;;) {
        if (threadIdx.x == 0) s_chunk = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t w0 = s_chunk * 2;
        if (w0 >= total) break;
        for (uint32_t w = w0; w < w0 + 2 && w < total; w++) {
            const uint32_t r = w >> 5, t = w & 31;
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(8192) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(tile)), "l"(in + ((size_t)r * 32 + t) * 8192), "r"(8192), "r"(bar_a) : "memory");
            }
            asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a), "r"(phase) : "memory");
            phase ^= 1;
            if (t < 24) {
                float4* dst = reinterpret_cast<float4*>(outA + ((size_t)r * 24 + t) * 32768);
                const float4 v = make_float4(1.f, 2.f, 3.f, (float)tile[threadIdx.x]);
                for (uint32_t g = threadIdx.x; g < 2048; g += blockDim.x)
                    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            } else {
                uint8_t* dst = outB + ((size_t)r * 8 + (t - 24)) * 327680;
                for (uint32_t c = warp; c < 128; c += 8, hit++) {
                    float* blk = hot + (hit & 1) * 640;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    blk[(lane * 20 + tile[c]) % 640] = 1.0f;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * 2560),
                                     "r"((uint32_t)__cvta_generic_to_shared(blk)), "r"(2560) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            __syncthreads();
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
extern "C" {
void mb_mix(const void* in, void* outA, void* outB, unsigned n_records, void* counter, int grid, void* stream) {
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaMemsetAsync(counter, 0, 4, (cudaStream_t)stream);
    k_mix<<<grid, 256, 8320 + 8 * 2 * 2560, (cudaStream_t)stream>>>((const uint8_t*)in, (uint8_t*)outA, (uint8_t*)outB, n_records,
                                                                   (unsigned int*)counter);
}
void mb_bulk_pattern(void* p, size_t nbytes, int grid, unsigned chunk, unsigned per_region, unsigned misalign, void* counter,
                     int depth, void* stream) {
    cudaFuncSetAttribute(k_bulk_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaMemsetAsync(counter, 0, 4, (cudaStream_t)stream);
    k_bulk_pattern<<<grid, 256, 8 * 2 * chunk, (cudaStream_t)stream>>>((uint8_t*)p, nbytes, chunk, per_region, misalign,
                                                                     (unsigned int*)counter, depth);
}
void mb_write(void* p, size_t nbytes, int grid, int block, int cs, void* stream) {
    k_write<<<grid, block, 0, (cudaStream_t)stream>>>((uint4*)p, nbytes / 16, cs);
}
void mb_read(const void* p, size_t nbytes, int grid, int block, void* sink, void* stream) {
    k_read<<<grid, block, 0, (cudaStream_t)stream>>>((const uint4*)p, nbytes / 16, (uint32_t*)sink);
}
void mb_bulk_write(void* p, size_t nbytes, int grid, unsigned chunk, void* stream) {
    cudaFuncSetAttribute(k_bulk_write, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_bulk_write<<<grid, 128, chunk, (cudaStream_t)stream>>>((uint8_t*)p, nbytes, chunk);
}
}
