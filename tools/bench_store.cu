// Developer microbenchmark (not part of the product path): how fast can the warps of ONE CTA per SM write a conv
// epilogue's output - scattered site rows of `row_bytes` contiguous bytes - with the store forms available?
//   mode 0  st.global.b32, lane = channel: one warp instruction = 128 contiguous bytes of one row (the weights-as-M epilogue)
//   mode 1  st.global.v4.b32, lane = (row, 16-byte piece): one instruction = 32 pieces of 32 different rows (sites-as-M, direct)
//   mode 2  st.global.v4.b32, 8 lanes = one 128-byte run (transposed through shared memory beforehand; the transpose is not timed)
//   mode 3  cp.async.bulk shared -> global, one bulk copy per row issued by lane 0 of each warp (the rows staged in shared memory)
//   mode 4  st.global.v2.b32, 4 lanes = one 32-byte sector of a row, 8 rows per instruction (the tcgen05.ld.16x256b register layout)
//   mode 5  st.global.v4.b32, 2 lanes = one 32-byte sector of a row, 16 rows per instruction
//   mode 6  st.global.v8.f32 (STG.256), lane = row: 32 full sectors of 32 different rows per instruction
//   mode 7  mode 0 with eight independent rows per loop iteration
// All 148 CTAs write disjoint regions of a 1 GB buffer; rows are visited in a hashed order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_store.bin tools/bench_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024, 1) k_store(char *buf, long long region, int row_bytes, int rows_per_warp, int mode, long long *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    char *base = buf + (long long)blockIdx.x * region;
    const uint32_t n_rows = (uint32_t)(region / row_bytes), row_mask = n_rows - 1u;        // a power of two
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((float *)sm)[i] = (float)i;
    __syncthreads();
    const long long t0 = clock64();
    const float v = (float)threadIdx.x;
    uint32_t h = 0x9e3779b9u * (uint32_t)(blockIdx.x * 64 + warp + 1);
    for (int r = 0; r < rows_per_warp; ++r) {
        h = h * 1664525u + 1013904223u;
        const uint32_t row = (h >> 8) & row_mask;
        char *dst = base + (size_t)row * (uint32_t)row_bytes;
        if (mode == 0) {
            for (int c = lane * 4; c < row_bytes; c += 128) *reinterpret_cast<float *>(dst + c) = v;
        } else if (mode == 1) {
            // 32 different rows per instruction: lane l writes piece j of row (row + l) for j = 0 .. row_bytes/16
            char *d2 = base + (size_t)((row + (uint32_t)lane) & row_mask) * (uint32_t)row_bytes;
            for (int c = 0; c < row_bytes; c += 16) *reinterpret_cast<float4 *>(d2 + c) = make_float4(v, v, v, v);
            r += 31;
        } else if (mode == 4) {
            // 8 different rows per instruction, 4 lanes = one full 32-byte sector of a row (the tcgen05.ld.16x256b register layout)
            char *d2 = base + (size_t)((row + (uint32_t)(lane >> 2)) & row_mask) * (uint32_t)row_bytes + (lane & 3) * 8;
            for (int c = 0; c < row_bytes; c += 32) *reinterpret_cast<float2 *>(d2 + c) = make_float2(v, v);
            r += 7;
        } else if (mode == 5) {
            // 16 different rows per instruction, 2 lanes = one 32-byte sector (st.v4)
            char *d2 = base + (size_t)((row + (uint32_t)(lane >> 1)) & row_mask) * (uint32_t)row_bytes + (lane & 1) * 16;
            for (int c = 0; c < row_bytes; c += 32) *reinterpret_cast<float4 *>(d2 + c) = make_float4(v, v, v, v);
            r += 15;
        } else if (mode == 6) {
            // 32 different rows per instruction, every lane one full 32-byte sector (st.global.v8.f32 = STG.256, sm_100)
            char *d2 = base + (size_t)((row + (uint32_t)lane) & row_mask) * (uint32_t)row_bytes;
            for (int c = 0; c < row_bytes; c += 32) asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(d2 + c), "f"(v) : "memory");
            r += 31;
        } else if (mode == 7) {
            // mode 0 with EIGHT independent rows per loop iteration (does a warp's rate depend on how many stores it has in flight?)
            uint32_t rows8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) rows8[k] = ((h >> 8) + 0x9e37u * (uint32_t)(k + 1) * 977u) & row_mask;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                char *d2 = base + (size_t)rows8[k] * (uint32_t)row_bytes;
                for (int c = lane * 4; c < row_bytes; c += 128) *reinterpret_cast<float *>(d2 + c) = v;
            }
            r += 7;
        } else if (mode == 2) {
            for (int c = lane * 16; c < row_bytes; c += 512) *reinterpret_cast<float4 *>(dst + c) = make_float4(v, v, v, v);
        } else {
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(sm + (warp & 7) * 2048)), "r"(row_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if ((r & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
            }
        }
    }
    if (mode == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    (void)nw;
}

int main()
{
    const long long region = 4LL << 20;        // 4 MB per CTA, 148 CTAs: 592 MB
    char *buf;
    long long *out;
    cudaMalloc(&buf, 148 * region);
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const char *names[8] = {"st.b32 lane=channel (128 B runs)", "st.v4 lane=row (32 rows/instr)", "st.v4 8+ lanes per row (512 B/instr)", "cp.async.bulk per row",
                            "st.v2 4 lanes = 32 B (8 rows/instr)", "st.v4 2 lanes = 32 B (16 rows/instr)", "st.v8 lane=row (32 rows x 32 B/instr)", "st.b32 lane=channel, 8 rows per iteration"};
    for (int row_bytes : {128, 256, 512})
        for (int mode = 0; mode < 8; ++mode)
            for (int warps : {4, 8, 16}) {
                const int rows_per_warp = 8192 * 4 / warps;
                for (int rep = 0; rep < 2; ++rep) {
                    k_store<<<148, warps * 32, 16384>>>(buf, region, row_bytes, rows_per_warp, mode, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                long long h[148];
                cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
                double avg = 0;
                for (int i = 0; i < 148; ++i) avg += (double)h[i];
                avg /= 148;
                const double bytes = (double)rows_per_warp * warps * row_bytes;
                printf("rows of %3d B  %-38s %2d warps: %6.1f B/cycle/SM  (%.0f cycles per row and SM)\n", row_bytes, names[mode], warps, bytes / avg,
                       avg / (rows_per_warp * warps));
            }
    return 0;
}
