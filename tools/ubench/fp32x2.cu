// Micro-benchmark: throughput of scalar FFMA/FADD/FMUL vs the sm_100 packed forms (FFMA2/FADD2/FMUL2).
// Decides whether the FFT butterflies should be written on float2 "lane pairs" (two independent lines per op).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu && ./fp32x2
// Reports FP instructions per clock per SMSP (clock64 inside the kernel, 16 resident warps per SMSP).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float a, float b, long long* cyc) {
    const long long c0 = clock64();
    float2 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, A.x, B.x); v[i].y = fmaf(v[i].y, A.y, B.y); }          // 2 FFMA
            if (MODE == 1) { v[i] = __ffma2_rn(v[i], A, B); }                                              // 1 FFMA2
            if (MODE == 2) { v[i].x = v[i].x + B.x; v[i].y = v[i].y + B.y; }                               // 2 FADD
            if (MODE == 3) { v[i] = __fadd2_rn(v[i], B); }                                                 // 1 FADD2
            if (MODE == 4) { v[i].x = v[i].x * A.x; v[i].y = v[i].y * A.y; }                               // 2 FMUL
            if (MODE == 5) { v[i] = __fmul2_rn(v[i], A); }                                                 // 1 FMUL2
            if (MODE == 6) { v[i] = __fadd2_rn(v[i], B); v[i] = __ffma2_rn(v[i], A, B); }                  // FADD2 + FFMA2
            if (MODE == 7) {                                                                               // butterfly-like: sum/diff of neighbours
                const float2 s = __fadd2_rn(v[i], v[(i + 1) % ILP]);
                v[i] = __ffma2_rn(s, A, make_float2(-v[i].x, -v[i].y));
            }
            if (MODE == 8) {                                                                               // same, scalar
                const float sx = v[i].x + v[(i + 1) % ILP].x, sy = v[i].y + v[(i + 1) % ILP].y;
                v[i].x = fmaf(sx, A.x, -v[i].x); v[i].y = fmaf(sy, A.y, -v[i].y);
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - c0;
}

template <int MODE>
void run(const char* name, int fp_instr_per_elem, int scalar_equiv_per_elem, float2* out, long long* cyc) {
    const int blocks = 148 * 8;   // 8 blocks x 8 warps per SM = 16 warps per SMSP, one wave
    for (int r = 0; r < 50; ++r) k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, cyc);   // warm-up: let the clocks ramp
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, cyc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 10;
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)cyc[i];
    avg /= blocks;
    const double per_smsp = 16.0 * ITER * ILP * fp_instr_per_elem / avg;
    printf("%-14s %7.3f ms  %5.2f FP instr/clk/SMSP  %5.2f scalar-equivalent/clk/SMSP  (SM clock ~%.2f GHz)\n", name, ms, per_smsp,
           per_smsp * scalar_equiv_per_elem / fp_instr_per_elem, avg / (ms * 1e6));
}

int main() {
    float2* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaMallocManaged(&cyc, 148 * 8 * sizeof(long long));
    run<0>("2xFFMA", 2, 2, out, cyc);
    run<1>("FFMA2", 1, 2, out, cyc);
    run<2>("2xFADD", 2, 2, out, cyc);
    run<3>("FADD2", 1, 2, out, cyc);
    run<4>("2xFMUL", 2, 2, out, cyc);
    run<5>("FMUL2", 1, 2, out, cyc);
    run<6>("FADD2+FFMA2", 2, 4, out, cyc);
    run<7>("bfly packed", 2, 4, out, cyc);
    run<8>("bfly scalar", 4, 4, out, cyc);
    cudaFree(out);
    return 0;
}
