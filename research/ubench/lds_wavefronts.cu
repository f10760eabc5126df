// Micro-benchmark (developer tool): cycles per warp-wide shared-memory load for the access patterns the track kernel could use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_wavefronts lds_wavefronts.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (double)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned off;   // in doubles
    if (MODE == 0) off = lane & 3;            // 4 rays per warp, LDS.64
    else if (MODE == 1) off = 0;              // uniform address, LDS.64
    else if (MODE == 2) off = lane & 1;       // 2 unique
    else if (MODE == 3) off = (lane >> 4);    // one unique address per half warp
    else if (MODE == 4) off = lane;           // 32 unique consecutive (no broadcast)
    else if (MODE == 5) off = 2 * (lane & 3); // LDS.128, 4 unique 16-byte words
    else if (MODE == 6) off = lane & 7;       // 8 unique
    else if (MODE == 7) off = ((lane >> 4) << 1) | (lane & 1);   // 2 unique per half warp, 4 per warp
    else if (MODE == 8) off = lane >> 3;       // 4 unique, one per quarter warp
    else if (MODE == 9) off = ((lane >> 3) & 1) | ((lane >> 4) << 1);   // quarter-warp uniform, pairs per half warp
    else off = 0;
    off = __shfl_sync(0xffffffffu, off, lane);
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int base = ((it & 3) * 32 + u) * 8;
            if (MODE == 5) {
                const double2 v = *reinterpret_cast<const double2*>(sm + base + off);
                acc0 += v.x; acc1 += v.y;
            } else {
                const double v = sm[base + off];
                if (u & 1) acc0 += v; else acc2 += v;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
    double* out; long long* cyc;
    const int blocks = 148 * 4, iters = 2000;
    cudaMalloc(&out, blocks * 256 * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<MODE><<<blocks, 256, 4096 * 8>>>(out, cyc, iters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256, 4096 * 8>>>(out, cyc, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SM: 4 CTAs x 8 warps x iters x 32 loads
    const double loads_per_sm = 4.0 * 8 * iters * 32;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-44s %.3f ms  %.2f SM-cycles per warp-load (at 1965 MHz; includes the 2-cycle DADD per load)\n", name, ms, cycles / loads_per_sm);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("LDS.64, 4 unique addresses (lane & 3)");
    run<1>("LDS.64, warp-uniform address");
    run<2>("LDS.64, 2 unique (lane & 1)");
    run<3>("LDS.64, one address per half warp");
    run<6>("LDS.64, 8 unique (lane & 7)");
    run<4>("LDS.64, 32 consecutive doubles");
    run<5>("LDS.128, 4 unique 16-byte words");
    run<7>("LDS.64, rays (0,1) in lanes 0-15, (2,3) in 16-31");
    run<8>("LDS.64, 4 unique, one per quarter warp");
    run<9>("LDS.64, 4 unique, quarter-warp uniform v2");
    return 0;
}
