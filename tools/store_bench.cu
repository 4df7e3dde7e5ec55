// Microbenchmark: how fast can 148 CTAs write the correlation block with the GEMM epilogue's access
// pattern?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/store_bench.cu -o store_bench
//   pattern 0: out[i][e][j] (row stride E*ld), a warp writes 32 rows x 128 B per chunk (current epilogue)
//   pattern 1: tile-major: every 256x128 tile of one epoch is one contiguous 128 KB block
//   pattern 2: like 0 but each warp writes 512 B per row (4 consecutive 128 B stores to the same row)
// warps per CTA: 8 or 16 (second argument).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void k_store(float *out, long nb, int E, long V2, long ld, int pattern, long tiles_j, long tiles_i)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long total = tiles_j * tiles_i * E;
    for (long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int e = (int)(tile / (tiles_j * tiles_i));
        const long rem = tile % (tiles_j * tiles_i);
        const long tj = rem / tiles_i, ti = rem % tiles_i;   // 128 cols x 256 rows per tile
        const float val = (float)tile;
        if (pattern == 0) {
            // 32 chunks of (32 rows x 32 cols): chunk = (q = col quarter, c = row chunk)
            for (int ch = warp; ch < 32; ch += nwarps) {
                const int q = ch & 3, c = ch >> 2;
                float *p = out + ((size_t)(ti * 256 + c * 32) * E + e) * ld + tj * 128 + q * 32 + lane;
                if (tj * 128 + q * 32 + lane < ld)
#pragma unroll
                    for (int r = 0; r < 32; r++) {
                        *p = val;
                        p += (size_t)E * ld;
                    }
            }
        } else if (pattern == 1) {
            float *p = out + (size_t)tile * 256 * 128;
            for (int ch = warp; ch < 32; ch += nwarps) {
#pragma unroll
                for (int r = 0; r < 32; r++) p[(size_t)(ch * 32 + r) * 32 + lane] = val;
            }
        } else {
            // each warp owns rows; writes the full 512 B of a row (4 x 128 B) before moving on
            for (int row = warp; row < 256; row += nwarps) {
                float *p = out + ((size_t)(ti * 256 + row) * E + e) * ld + tj * 128 + lane;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (tj * 128 + q * 32 + lane < ld) p[q * 32] = val;
            }
        }
    }
}

int main(int argc, char **argv)
{
    const long nb = 2048, V2 = 50000, ld = 50016;
    const int E = 32;
    float *out;
    const long tiles_j = (V2 + 127) / 128, tiles_i = nb / 256;
    size_t bytes = (size_t)tiles_j * tiles_i * E * 256 * 128 * 4 + (size_t)nb * E * 64 * 4;
    if (cudaMalloc(&out, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    for (int pattern = 0; pattern < 3; pattern++)
        for (int warps = 8; warps <= 32; warps *= 2) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            for (int it = 0; it < 2; it++) k_store<<<148, warps * 32>>>(out, nb, E, V2, ld, pattern, tiles_j, tiles_i);
            cudaEventRecord(e0);
            const int reps = 5;
            for (int it = 0; it < reps; it++) k_store<<<148, warps * 32>>>(out, nb, E, V2, ld, pattern, tiles_j, tiles_i);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("pattern %d warps %2d : %.3f ms  %.0f GB/s  (%s)\n", pattern, warps, ms,
                   (double)nb * E * V2 * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    // grid = 296 (2 CTAs per SM) with pattern 0
    for (int g = 296; g <= 592; g *= 2) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_store<<<g, 256>>>(out, nb, E, V2, ld, 0, tiles_j, tiles_i);
        cudaEventRecord(e0);
        for (int it = 0; it < 5; it++) k_store<<<g, 256>>>(out, nb, E, V2, ld, 0, tiles_j, tiles_i);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        printf("pattern 0 warps  8 grid %d : %.3f ms  %.0f GB/s\n", g, ms, (double)nb * E * V2 * 4 / ms / 1e6);
    }
    return 0;
}
