// Developer microbenchmark (not part of the product path): how fast does ONE CTA per SM stream a weight image from L2 into
// shared memory, the way the tcgen05 conv kernels' loader does it?
//   mode 0  cp.async.bulk global -> shared, one elected thread, `depth` copies of `chunk` bytes in flight (a ring of mbarriers)
//   mode 1  the same bytes by cp.async.cg (LDGSTS, 16 bytes per thread) from `warps` warps, `depth` chunk-sized groups in flight
// The image (1.2 MB, conv5's weight tile) is re-read `passes` times: it stays in L2.  grid = 1, 4, 148 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_bulk.bin tools/bench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_bulk(const char *img, long long img_bytes, int chunk, int depth, int passes, int mode, long long *out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[8];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_chunks = img_bytes / chunk * passes;
    const long long per_pass = img_bytes / chunk;
    const long long t0 = clock64();
    if (mode == 0) {
        if (tid == 0) {
            for (long long q = 0; q < n_chunks + depth; ++q) {
                if (q >= depth) mbar_wait(smem_u32(&bar[(q - depth) % depth]), (uint32_t)(((q - depth) / depth) & 1));    // chunk q - depth has landed: its slot is free
                if (q < n_chunks) {
                    const uint32_t b = smem_u32(&bar[q % depth]);
                    mbar_expect_tx(b, (uint32_t)chunk);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (q % depth) * chunk)),
                                 "l"(img + (q % per_pass) * chunk), "r"((uint32_t)chunk), "r"(b)
                                 : "memory");
                }
            }
        }
    } else if (mode >= 2) {
        // mode 2 / 3: TWO / FOUR issuing threads (lane 0 of different warps), each with its own barriers and slots, chunks round-robin
        // mode 4 / 5: two / four LANES of one warp, converged (one warp instruction issues their copies together)
        // mode 6 / 7: two / four lanes of one warp in DIVERGENT branches (each lane runs the loop on its own path)
        const int ni = (mode == 2 || mode == 4 || mode == 6) ? 2 : 4;
        const bool lanes = mode >= 4;
        const int w = lanes ? tid : tid >> 5;
        auto body = [&]() {
            const long long mine = n_chunks / ni;
            for (long long q = 0; q < mine + depth; ++q) {
                if (q >= depth) mbar_wait(smem_u32(&bar[w * 2 + (q - depth) % depth]), (uint32_t)(((q - depth) / depth) & 1));
                if (q < mine) {
                    const uint32_t b = smem_u32(&bar[w * 2 + q % depth]);
                    mbar_expect_tx(b, (uint32_t)chunk);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + ((size_t)w * depth + q % depth) * chunk)),
                                 "l"(img + ((q * ni + w) % per_pass) * chunk), "r"((uint32_t)chunk), "r"(b)
                                 : "memory");
                }
            }
        };
        if (mode >= 6) {
            // four textually separate call sites: the lanes are on different paths
            if (tid == 0) body();
            else if (tid == 1) body();
            else if (tid == 2 && ni == 4) body();
            else if (tid == 3 && ni == 4) body();
        } else if (lanes ? tid < ni : ((tid & 31) == 0 && w < ni)) {
            body();
        }
    } else {
        // every thread copies 16 bytes of each 16 * blockDim-byte slice; groups of one chunk, `depth` of them in flight
        const int slices = chunk / (16 * (int)blockDim.x);
        for (long long q = 0; q < n_chunks + depth; ++q) {
            if (q < n_chunks) {
                const char *src = img + (q % per_pass) * chunk;
                const uint32_t dst = smem_u32(sm + (q % depth) * chunk);
                for (int s = 0; s < slices; ++s) {
                    const int off = (s * (int)blockDim.x + tid) * 16;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(src + off) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (q >= depth - 1) {
                switch (depth) {
                case 1: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                case 3: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main()
{
    const long long img_bytes = 36LL * 32768;      // conv5: 36 K blocks x 32 KB (one 128-row weight tile, hi + lo)
    char *img;
    long long *out;
    cudaMalloc(&img, img_bytes);
    cudaMemset(img, 1, img_bytes);
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int passes = 8;
    for (int grid : {148})
        for (int mode : {0, 4, 6, 7})
            for (int chunk : {8192, 16384, 32768})
                for (int depth : {1, 2}) {
                    const int ni = mode == 0 ? 1 : (mode == 2 || mode == 4 || mode == 6) ? 2 : 4;
                    if ((long long)chunk * depth * ni > 192 * 1024) continue;
                    for (int rep = 0; rep < 2; ++rep) {
                        k_bulk<<<grid, 128, (size_t)chunk * depth * ni>>>(img, img_bytes, chunk, depth, passes, mode, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    long long h[148];
                    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double avg = 0;
                    for (int i = 0; i < grid; ++i) avg += (double)h[i];
                    avg /= grid;
                    printf("grid %3d  cp.async.bulk from %d %s  chunk %5d B  depth %d per thread: %6.1f B/cycle/SM  (%.0f cycles per chunk)\n", grid, ni, mode >= 6 ? "DIVERGENT lanes of one warp" : mode >= 4 ? "lanes of one warp" : "thread(s) (warps) ", chunk, depth,
                           (double)img_bytes * passes / avg, avg / ((double)img_bytes / chunk * passes));
                }
    return 0;
}
