// trace.cu — per-CTA timeline of the row kernel (development tool). Build with -DOW_TRACE:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DOW_TRACE -I../../fft-ocean-waves_b200/csrc -o trace trace.cu
//   ./trace [N]      prints, per phase, the distribution of durations (SM cycles) over CTAs and the launch's overall timeline.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ow_frame_kernels.cuh"

using namespace ow;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_h0(float4* h0, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { unsigned s = (unsigned)i * 747796405u + 12345u; auto r = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; }; h0[i] = make_float4(r(), r(), r(), r()); }
}
__global__ void fill_ktab(float* k, int N, float L) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < N) k[i] = (2.0f * 3.14159265f * ((float)i - N / 2.0f)) / L; }

template <int N>
void run() {
    using C = Cfg<N>; using R = typename C::Row;
    const size_t nn = (size_t)N * N;
    float4* h0; float* ktab; float2* inter; unsigned long long* tr;
    CK(cudaMalloc(&h0, nn * 16)); CK(cudaMalloc(&ktab, N * 4)); CK(cudaMalloc(&inter, nn / 2 * 3 * 8 * 4));
    const int ncta = N / 2 / C::ROW_PAIRS;
    CK(cudaMalloc(&tr, (size_t)ncta * 8 * 8));
    CK(cudaMemcpyToSymbol(g_ow_trace, &tr, sizeof(tr)));
    fill_h0<<<(unsigned)((nn + 255) / 256), 256>>>(h0, nn);
    fill_ktab<<<(N + 255) / 256, 256>>>(ktab, N, 1000.0f);
    float4 *hp, *nyq;
    CK(cudaMalloc(&hp, nn / 2 * 24)); CK(cudaMalloc(&nyq, (size_t)(N / 2) * 16));
    fill_h0<<<(unsigned)((nn / 2 * 3 / 2 + 255) / 256), 256>>>(hp, nn / 2 * 3 / 2);
    fill_h0<<<(unsigned)((N / 2 + 255) / 256), 256>>>(nyq, N / 2);
    FrameBuffers fb{N, h0, hp, nyq, ktab, nullptr, inter, nullptr, nullptr, nullptr, 0};
    SlotTable tab{};
    auto k = ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, C::ROW_PAIRS>()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int r = 0; r < 5; ++r) {
        tab.slot[0] = r % 4; tab.time[0] = 1.0f + r;
        cudaEventRecord(e0);
        k<<<dim3(ncta, 1), R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>()>>>(fb, tab);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
    }
    std::vector<unsigned long long> h((size_t)ncta * 8);
    CK(cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost));
    printf("N=%d row kernel: %d CTAs x %d threads, last launch %.2f us\n", N, ncta, R::T * C::ROW_PAIRS, ms * 1e3);
    const char* names[5] = {"phase0 (loads+spectrum+stage0)", "barrier 1", "phase1 (stage 1)", "barrier 2", "phase2 (stage 2 + stores)"};
    for (int ph = 0; ph < 5; ++ph) {
        std::vector<double> d;
        for (int c = 0; c < ncta; ++c) d.push_back((double)(h[c * 8 + ph + 2] - h[c * 8 + ph + 1]));
        std::sort(d.begin(), d.end());
        printf("  %-32s cycles: min %7.0f  p10 %7.0f  median %7.0f  p90 %7.0f  max %7.0f\n", names[ph], d[0], d[d.size() / 10], d[d.size() / 2], d[d.size() * 9 / 10], d.back());
    }
    std::vector<double> tot, start;
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < ncta; ++c) t0 = std::min(t0, h[c * 8 + 7]);
    for (int c = 0; c < ncta; ++c) { tot.push_back((double)(h[c * 8 + 6] - h[c * 8 + 1])); start.push_back((double)(h[c * 8 + 7] - t0)); }
    std::vector<double> ts = tot; std::sort(ts.begin(), ts.end());
    printf("  CTA lifetime cycles: min %.0f median %.0f p90 %.0f max %.0f  (%.2f us median at 1.965 GHz)\n", ts[0], ts[ts.size() / 2], ts[ts.size() * 9 / 10], ts.back(), ts[ts.size() / 2] / 1965.0);
    // start-time histogram (globaltimer ns)
    std::vector<double> ss = start; std::sort(ss.begin(), ss.end());
    printf("  CTA start times (us after the first): p25 %.2f  p50 %.2f  p75 %.2f  p90 %.2f  max %.2f\n", ss[ss.size() / 4] / 1e3, ss[ss.size() / 2] / 1e3, ss[ss.size() * 3 / 4] / 1e3, ss[ss.size() * 9 / 10] / 1e3, ss.back() / 1e3);
    // CTAs per SM
    int per_sm[256] = {0};
    for (int c = 0; c < ncta; ++c) per_sm[h[c * 8] & 0xff]++;
    int mn = 1 << 30, mx = 0; for (int i = 0; i < 148; ++i) { mn = std::min(mn, per_sm[i]); mx = std::max(mx, per_sm[i]); }
    printf("  CTAs per SM: min %d max %d\n", mn, mx);
}

// Column kernel timeline: `count` slots per launch (steady state: several waves of tiles per SM).
template <int N>
void run_col(int count) {
    using C = Cfg<N>; using K = typename C::Col;
    const size_t nn = (size_t)N * N;
    float2* inter; float* disp; unsigned long long* tr;
    CK(cudaMalloc(&inter, nn / 2 * 3 * 8 * count)); CK(cudaMalloc(&disp, nn * 3 * 4 * count));
    fill_h0<<<(unsigned)((nn / 2 * 3 * count / 2 + 255) / 256), 256>>>(reinterpret_cast<float4*>(inter), nn / 2 * 3 * count / 2);
    const int ncta = N / (2 * C::COL_G) * 3 * count;
    CK(cudaMalloc(&tr, (size_t)ncta * 8 * 8));
    CK(cudaMemcpyToSymbol(g_ow_trace, &tr, sizeof(tr)));
    FrameBuffers fb{}; fb.N = N; fb.inter = inter; fb.disp = disp;
    SlotTable tab{};
    for (int i = 0; i < count; ++i) tab.slot[i] = i;
    auto k = ow_col_kernel<K, C::COL_G, C::COL_MINB>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, C::COL_G>::SMEM));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<<<dim3(N / (2 * C::COL_G), 3, count), K::T * C::COL_G, ColLayout<K, C::COL_G>::SMEM>>>(fb, tab, 1.0f);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
    }
    std::vector<unsigned long long> h((size_t)ncta * 8);
    CK(cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost));
    printf("N=%d column kernel: %d CTAs x %d threads (%d slots), last launch %.2f us = %.2f us/frame\n", N, ncta, K::T * C::COL_G, count, ms * 1e3, ms * 1e3 / count);
    const char* names[5] = {"phase0 (loads+unpack+stage0)", "barrier 1", "phase1 (stage 1)", "barrier 2", "phase2 (stage 2 + stores)"};
    for (int ph = 0; ph < 5; ++ph) {
        std::vector<double> d;
        for (int c = 0; c < ncta; ++c) d.push_back((double)(h[c * 8 + ph + 2] - h[c * 8 + ph + 1]));
        std::sort(d.begin(), d.end());
        printf("  %-32s cycles: min %7.0f  p10 %7.0f  median %7.0f  p90 %7.0f  max %7.0f\n", names[ph], d[0], d[d.size() / 10], d[d.size() / 2], d[d.size() * 9 / 10], d.back());
    }
    std::vector<double> tot;
    for (int c = 0; c < ncta; ++c) tot.push_back((double)(h[c * 8 + 6] - h[c * 8 + 1]));
    std::sort(tot.begin(), tot.end());
    printf("  CTA lifetime cycles (thread 0): min %.0f median %.0f p90 %.0f max %.0f  (%.2f us median at 1.965 GHz)\n", tot[0], tot[tot.size() / 2], tot[tot.size() * 9 / 10], tot.back(), tot[tot.size() / 2] / 1965.0);
    // gap between consecutive CTAs on one SM: sort the CTAs of each SM by start time (globaltimer), gap = next start - this end is not
    // available in one clock domain; report CTAs per SM and launch duration / CTAs per SM instead
    int per_sm[256] = {0};
    for (int c = 0; c < ncta; ++c) per_sm[h[c * 8] & 0xff]++;
    int mn = 1 << 30, mx = 0; for (int i = 0; i < 148; ++i) { mn = std::min(mn, per_sm[i]); mx = std::max(mx, per_sm[i]); }
    printf("  CTAs per SM: min %d max %d -> %.0f cycles of launch time per CTA slot\n", mn, mx, ms * 1e-3 * 1.965e9 / mx);
    cudaFree(inter); cudaFree(disp); cudaFree(tr);
}

int main(int argc, char** argv) {
    if (argc > 2 && argv[2][0] == 'c') {
        const int N = atoi(argv[1]), count = argc > 3 ? atoi(argv[3]) : 8;
        switch (N) { case 512: run_col<512>(count); break; case 1024: run_col<1024>(count); break; case 2048: run_col<2048>(count); break; default: printf("N?\n"); }
        return 0;
    }
    const int N = argc > 1 ? atoi(argv[1]) : 2048;
    switch (N) { case 512: run<512>(); break; case 1024: run<1024>(); break; case 2048: run<2048>(); break; case 4096: run<4096>(); break; default: printf("N?\n"); }
    return 0;
}
