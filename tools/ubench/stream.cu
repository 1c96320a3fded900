// Micro-benchmark: what a plain streaming kernel achieves at the sizes of ONE N=2048 frame kernel (tens of MB,
// 20-35 us launches), so the frame kernels are judged against an achievable number and not only the 2 GiB copy peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream stream.cu && ./stream
// Patterns (rotating over 8 buffer sets, like consecutive frames of a sweep):
//   row-like   : read 16 B/texel, write 12 B/texel
//   col-like   : read 12 B/texel (just written by the previous kernel -> L2), write 12 B/texel
//   normal-like: read 4 B/texel, write 16 B/texel
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_stream(const float4* __restrict__ in, size_t n_in, float4* __restrict__ out, size_t n_out) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = gridDim.x * (size_t)blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = tid; i < n_in; i += nt) {
        const float4 v = __ldg(in + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (size_t i = tid; i < n_out; i += nt) out[i] = acc;
}

int main() {
    const size_t texels = 2048ull * 2048ull;
    const int sets = 8;
    float4 *a, *b, *c;
    cudaMalloc(&a, texels * 16);                  // h0-like input (re-read every frame)
    cudaMalloc(&b, texels * 16 * sets);
    cudaMalloc(&c, texels * 16 * sets);
    cudaMemset(a, 0, texels * 16);
    cudaEvent_t e[4];
    for (auto& x : e) cudaEventCreate(&x);
    for (int grid : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        float us[3] = {0, 0, 0};
        const int reps = 20;
        for (int r = -2; r < reps; ++r) {
            const int s = (r + 2) % sets;
            float4* inter = b + (size_t)s * texels;         // 12 B/texel = 0.75 float4 per texel
            float4* disp = c + (size_t)s * texels;
            float4* nrm = b + (size_t)(s ^ 1) * texels;
            cudaEventRecord(e[0]);
            k_stream<<<grid, 256>>>(a, texels, inter, texels * 3 / 4);
            cudaEventRecord(e[1]);
            k_stream<<<grid, 256>>>(inter, texels * 3 / 4, disp, texels * 3 / 4);
            cudaEventRecord(e[2]);
            k_stream<<<grid, 256>>>(disp, texels / 4, nrm, texels);
            cudaEventRecord(e[3]);
            cudaEventSynchronize(e[3]);
            if (r < 0) continue;
            for (int i = 0; i < 3; ++i) { float ms; cudaEventElapsedTime(&ms, e[i], e[i + 1]); us[i] += ms * 1e3f / reps; }
        }
        const double mb[3] = {28.0 * texels / 1e6, 24.0 * texels / 1e6, 20.0 * texels / 1e6};
        printf("grid %4d: row-like %6.2f us (%5.0f GB/s)  col-like %6.2f us (%5.0f GB/s)  normal-like %6.2f us (%5.0f GB/s)\n", grid,
               us[0], mb[0] / us[0] * 1e3, us[1], mb[1] / us[1] * 1e3, us[2], mb[2] / us[2] * 1e3);
    }
    return 0;
}
