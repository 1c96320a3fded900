// tune.cu — variant sweeper for the three frame kernels (development tool, not part of the product).
// Instantiates the library's own kernel templates (csrc/ow_frame_kernels.cuh) with alternative template
// parameters and times each with CUDA events inside the real row -> column -> normal sequence (so the column
// kernel sees the intermediate in L2 the way the product does). Repetitions rotate over several output sets, so the
// caches are in the steady state of a multi-frame sweep (previous frames' outputs draining from L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../fft-ocean-waves_b200/csrc -o tune tune.cu
//   ./tune N count [reps]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "ow_frame_kernels.cuh"

using namespace ow;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Ctx {
    int N, count, reps, nsets;
    FrameBuffers fb;
    SlotTable tab[8];      // rotating output sets: rep r writes set r % nsets, like consecutive frames of a sweep
    mutable int cur = 0;
    mutable cudaStream_t st = nullptr;   // stream the next launches go to
    cudaEvent_t ev[4];
};

__global__ void fill_h0(float4* h0, size_t n, unsigned seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned s = (unsigned)i * 747796405u + seed;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    h0[i] = make_float4(rnd(), rnd(), rnd(), rnd());
}
__global__ void fill_ktab(float* k, int N, float L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) k[i] = (2.0f * 3.14159265f * ((float)i - N / 2.0f)) / L;
}

using Launch = std::function<void(const Ctx&)>;

template <int N> Launch default_row() {
    using C = Cfg<N>; using R = typename C::Row;
    auto k = ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, C::ROW_PAIRS>()));
    return [k](const Ctx& c) { k<<<dim3(N / 2 / C::ROW_PAIRS, c.count), R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>(), c.st>>>(c.fb, c.tab[c.cur]); };
}
template <class R, int PAIRS, int MINB> Launch row_variant() {
    auto k = ow_row_kernel<R, PAIRS, MINB, true>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, PAIRS>()));
    return [k](const Ctx& c) { k<<<dim3(R::N / 2 / PAIRS, c.count), R::T * PAIRS, row_smem<R, PAIRS>(), c.st>>>(c.fb, c.tab[c.cur]); };
}
// persistent pipelined row kernel: grid = min(items, SMs x resident CTAs)
template <class R, int PAIRS, int MINB> Launch row_pipe_variant() {
    auto k = ow_row_pipe_kernel<R, PAIRS, MINB, true>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, PAIRS>()));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, R::T * PAIRS, row_smem<R, PAIRS>()));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    printf("    [row pipe PAIRS=%d minb=%d: %d regs, %d CTAs/SM, %zu B local]\n", PAIRS, MINB, fa.numRegs, per_sm, fa.localSizeBytes);
    return [k, per_sm](const Ctx& c) {
        const int items = c.count * (R::N / 2 / PAIRS);
        const int grid = items < 148 * per_sm ? items : 148 * per_sm;
        k<<<grid, R::T * PAIRS, row_smem<R, PAIRS>(), c.st>>>(c.fb, c.tab[c.cur], items);
    };
}
template <int N> Launch default_col() {
    using C = Cfg<N>; using K = typename C::Col;
    auto k = ow_col_kernel<K, C::COL_G, C::COL_MINB>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, C::COL_G>::SMEM));
    return [k](const Ctx& c) {
        k<<<dim3(N / (2 * C::COL_G), 3, c.count), K::T * C::COL_G, ColLayout<K, C::COL_G>::SMEM, c.st>>>(c.fb, c.tab[c.cur], 0.5f / ((float)N * N));
    };
}
template <class K, int G, int MINB> Launch col_variant() {
    auto k = ow_col_kernel<K, G, MINB>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G>::SMEM));
    return [k](const Ctx& c) {
        k<<<dim3(K::N / (2 * G), 3, c.count), K::T * G, ColLayout<K, G>::SMEM, c.st>>>(c.fb, c.tab[c.cur], 0.5f / ((float)K::N * K::N));
    };
}
template <int N, bool JAC, int RY, int WARPS, int MINB> Launch nrm_variant() {
    auto k = ow_normal_kernel<N, JAC, RY, WARPS, MINB>;
    return [k](const Ctx& c) { k<<<dim3(N / 128, N / (WARPS * RY), c.count), dim3(32, WARPS), 0, c.st>>>(c.fb, c.tab[c.cur]); };
}

// Runs the sequence `reps` times and returns the mean duration (us) of each of the three kernels.
void time_seq(const Ctx& c, const Launch& row, const Launch& col, const Launch& nrm, float us[3]) {
    us[0] = us[1] = us[2] = 0;
    for (int r = -2; r < c.reps; ++r) {
        c.cur = (r + 2) % c.nsets;
        CK(cudaEventRecord(c.ev[0]));
        row(c);
        CK(cudaEventRecord(c.ev[1]));
        col(c);
        CK(cudaEventRecord(c.ev[2]));
        nrm(c);
        CK(cudaEventRecord(c.ev[3]));
        CK(cudaEventSynchronize(c.ev[3]));
        CK(cudaGetLastError());
        if (r < 0) continue;
        for (int i = 0; i < 3; ++i) { float ms; CK(cudaEventElapsedTime(&ms, c.ev[i], c.ev[i + 1])); us[i] += ms * 1e3f / c.reps; }
    }
}

// Throughput mode, like ow_step_multi: the output sets are independent frame groups, spread round-robin over 3 streams.
double time_sweep(const Ctx& c, const Launch& row, const Launch& col, const Launch& nrm) {
    static cudaStream_t ss[3] = {nullptr, nullptr, nullptr};
    if (!ss[0]) for (auto& s : ss) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    double best = 1e30;
    for (int r = -1; r < 3; ++r) {
        CK(cudaDeviceSynchronize());
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        CK(cudaEventRecord(a, ss[0]));
        CK(cudaStreamWaitEvent(ss[1], a, 0)); CK(cudaStreamWaitEvent(ss[2], a, 0));
        const int groups = 8 * c.nsets;
        for (int g = 0; g < groups; ++g) {
            c.cur = g % c.nsets; c.st = ss[g % 3];
            row(c); col(c); nrm(c);
        }
        cudaEvent_t j1, j2; cudaEventCreate(&j1); cudaEventCreate(&j2);
        CK(cudaEventRecord(j1, ss[1])); CK(cudaEventRecord(j2, ss[2]));
        CK(cudaStreamWaitEvent(ss[0], j1, 0)); CK(cudaStreamWaitEvent(ss[0], j2, 0));
        CK(cudaEventRecord(b, ss[0]));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (r >= 0) best = std::min(best, (double)ms * 1e3 / (groups * c.count));
        cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(j1); cudaEventDestroy(j2);
    }
    c.st = nullptr;
    return best;   // us per frame
}

#ifdef TUNE_SWEEP
template <int N>
void sweep(Ctx& c) {
    using C = Cfg<N>; using R = typename C::Row; using K = typename C::Col;
    const Launch rowc = default_row<N>(), col0 = default_col<N>();
    const Launch nrm0 = nrm_variant<N, true, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB>();
    const Launch nrm1 = nrm_variant<N, false, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB>();
    const Launch none = [](const Ctx&) {};
    auto rep = [&](const char* what, const Launch& r, const Launch& k, const Launch& n) {
        printf("%-54s %8.2f us/frame\n", what, time_sweep(c, r, k, n)); fflush(stdout);
    };
    const Launch rowp = row_pipe_variant<R, C::ROW_PAIRS, C::ROW_MINB>();
    rep("all: classic row, default col, normal+J", rowc, col0, nrm0);
    rep("all: pipe row,    default col, normal+J", rowp, col0, nrm0);
    rep("all: pipe row,    default col, normal (no J)", rowp, col0, nrm1);
#ifdef TUNE_SWEEP_BRIEF
    rep("col only: default", none, col0, none);
    return;
#endif
    rep("row only: classic", rowc, none, none);
    rep("row only: pipe (config)", rowp, none, none);
#define RP(PAIRS, MB) rep("row only: pipe PAIRS=" #PAIRS " minb=" #MB, row_pipe_variant<R, PAIRS, MB>(), none, none);
#define RC(PAIRS, MB) rep("row only: classic PAIRS=" #PAIRS " minb=" #MB, row_variant<R, PAIRS, MB>(), none, none);
    RP(1, 2) RP(1, 3) RP(1, 4) RP(1, 5) RP(2, 2) RP(2, 3)
    RC(1, 3) RC(1, 4) RC(1, 5) RC(1, 6) RC(2, 2) RC(2, 3)
    if (N <= 1024) { RP(4, 1) RP(4, 2) RP(4, 3) RP(4, 4) RC(4, 3) RC(4, 4) }
    {   // twice the threads per line (half the work per thread)
        using R2x = Plan<N, R::R0, R::R1, R::R2, 2 * R::T, R::S1 - R::R2, R::S0 - R::R1 * R::S1>;
#define RC2(PAIRS, MB) rep("row only: classic 2xT PAIRS=" #PAIRS " minb=" #MB, row_variant<R2x, PAIRS, MB>(), none, none);
#define RP2(PAIRS, MB) rep("row only: pipe 2xT PAIRS=" #PAIRS " minb=" #MB, row_pipe_variant<R2x, PAIRS, MB>(), none, none);
        RC2(1, 2) RC2(1, 3) RC2(1, 4) RC2(2, 2) RC2(2, 3)
        if constexpr (R::M % (2 * R::T) == 0) { RP2(1, 2) RP2(1, 3) RP2(2, 2) }      // the pipelined kernel needs whole stage-0 batches
        if (N <= 1024) { RC2(4, 1) RC2(4, 2) RC2(1, 8) RC2(2, 4) }
        using K2x = Plan<N, K::R0, K::R1, K::R2, 2 * K::T, K::S1 - K::R2, K::S0 - K::R1 * K::S1>;
        if (N == 2048) {
            rep("all+J: classic cfg row, default col, normal+J", rowc, col0, nrm0);
            rep("all+J: classic cfg row, G4 minb1 col, normal+J", rowc, col_variant<K, 4, 1>(), nrm0);
            rep("all+J: classic cfg row, G4 minb2 col, normal+J", rowc, col_variant<K, 4, 2>(), nrm0);
            rep("all+J: classic cfg row, G4 minb3 col, normal+J", rowc, col_variant<K, 4, 3>(), nrm0);
        }
        if (N <= 1024) {
            rep("all: classic 2xT P2 minb2 row, default col, normal", row_variant<R2x, 2, 2>(), col0, nrm1);
            rep("all: classic 2xT P2 minb4 row, default col, normal", row_variant<R2x, 2, 4>(), col0, nrm1);
            rep("all: classic cfg row, G4 minb2 col, normal", rowc, col_variant<K, 4, 2>(), nrm1);
            rep("all: classic cfg row, G4 minb4 col, normal", rowc, col_variant<K, 4, 4>(), nrm1);
            rep("all: classic cfg row, G8 minb3 col, normal", rowc, col_variant<K, 8, 3>(), nrm1);
            rep("all: classic cfg row, G8 minb4 col, normal", rowc, col_variant<K, 8, 4>(), nrm1);
            rep("all: classic 2xT P2 minb4 row, G8 minb3 col, normal", row_variant<R2x, 2, 4>(), col_variant<K, 8, 3>(), nrm1);
        }
        rep("all: classic 2xT P1 minb3 row, default col, normal", row_variant<R2x, 1, 3>(), col0, nrm1);
        rep("all: classic P1 minb4 row, default col, normal", row_variant<R, 1, 4>(), col0, nrm1);
        rep("all: classic cfg row, default col, normal", rowc, col0, nrm1);
        rep("all: pipe cfg row, default col, normal", rowp, col0, nrm1);
        rep("all: pipe P1 minb4 row, default col, normal", row_pipe_variant<R, 1, 4>(), col0, nrm1);
        rep("all: pipe cfg row, 2xT G8 col, normal", rowp, col_variant<K2x, 8, 1>(), nrm1);
        rep("all: classic P1 minb4 row, 2xT G8 col, normal", row_variant<R, 1, 4>(), col_variant<K2x, 8, 1>(), nrm1);
    }
    rep("col only: default", none, col0, none);
#define CV(G, MB) rep("col only: G=" #G " minb=" #MB, none, col_variant<K, G, MB>(), none);
    CV(8, 1) CV(8, 2) CV(4, 1) CV(4, 2) CV(4, 3) CV(4, 4)
    {
        using K2x = Plan<N, K::R0, K::R1, K::R2, 2 * K::T, K::S1 - K::R2, K::S0 - K::R1 * K::S1>;
#define CV2(G, MB) rep("col only: 2xT G=" #G " minb=" #MB, none, col_variant<K2x, G, MB>(), none);
        CV2(8, 1) CV2(4, 1) CV2(4, 2)
    }
    rep("normal+J only", none, none, nrm0);
    rep("normal only", none, none, nrm1);
}
#else
template <int N>
void sweep(Ctx& c) {
    const Launch row0 = default_row<N>(), col0 = default_col<N>();
    const Launch nrm0 = nrm_variant<N, true, Cfg<N>::NRM_RY, Cfg<N>::NRM_WARPS, Cfg<N>::NRM_MINB>();
    float us[3];
    const double texels = (double)N * N * c.count;
    auto report = [&](const char* what, int which) {
        static const double bytes[3] = {28, 24, 32};
        printf("%-44s %8.2f us  (%6.0f GB/s on %g B/texel)   [row %.1f col %.1f nrm %.1f]\n", what, us[which],
               bytes[which] * texels / us[which] / 1e3, bytes[which], us[0], us[1], us[2]);
        fflush(stdout);
    };
    time_seq(c, row0, col0, nrm0, us);
    report("defaults: row", 0); report("defaults: col", 1); report("defaults: normal+J", 2);
#ifdef TUNE_MINIMAL
    return;
#endif
#ifdef TUNE_OCC
    {   // same radices and paddings, twice the threads per line (half the work per thread), register caps via MINB
        using R0_ = typename Cfg<N>::Row;
        using R2x = Plan<N, R0_::R0, R0_::R1, R0_::R2, 2 * R0_::T, R0_::S1 - R0_::R2, R0_::S0 - R0_::R1 * R0_::S1>;
        constexpr int PR = Cfg<N>::ROW_PAIRS;
#define ROW2(PAIRS, MB) { time_seq(c, row_variant<R2x, PAIRS, MB>(), col0, nrm0, us); report("row 2xT PAIRS=" #PAIRS " minb=" #MB, 0); }
        ROW2(1, 1) ROW2(1, 2) ROW2(1, 3) ROW2(1, 4) ROW2(1, 6) ROW2(1, 8) ROW2(2, 2) ROW2(2, 3) ROW2(2, 4)
        (void)PR;
        using K0_ = typename Cfg<N>::Col;
        using K2x = Plan<N, K0_::R0, K0_::R1, K0_::R2, 2 * K0_::T, K0_::S1 - K0_::R2, K0_::S0 - K0_::R1 * K0_::S1>;
#define COL2(G, MB) { time_seq(c, row0, col_variant<K2x, G, MB>(), nrm0, us); report("col 2xT G=" #G " minb=" #MB, 1); }
        COL2(4, 1) COL2(4, 2) COL2(4, 3) COL2(8, 1) COL2(8, 2) COL2(2, 4)
    }
    return;
#endif
#define NRM(RY, W, MB) { time_seq(c, row0, col0, nrm_variant<N, true, RY, W, MB>(), us); report("normal+J RY=" #RY " warps=" #W " minb=" #MB, 2); }
    NRM(4, 4, 4) NRM(8, 4, 4) NRM(16, 4, 4) NRM(8, 8, 2) NRM(16, 8, 2) NRM(8, 2, 8) NRM(16, 2, 8) NRM(4, 8, 2) NRM(8, 4, 5) NRM(8, 4, 3)
#define NRM0(RY, W, MB) { time_seq(c, row0, col0, nrm_variant<N, false, RY, W, MB>(), us); report("normal    RY=" #RY " warps=" #W " minb=" #MB, 2); }
    NRM0(8, 4, 4) NRM0(16, 4, 4) NRM0(8, 8, 2) NRM0(8, 4, 6)
    using C = Cfg<N>;
    using R = typename C::Row;
    using K = typename C::Col;
#define ROWV(PAIRS, MB) { time_seq(c, row_variant<R, PAIRS, MB>(), col0, nrm0, us); report("row PAIRS=" #PAIRS " minb=" #MB, 0); }
    ROWV(1, 1) ROWV(1, 2) ROWV(1, 3) ROWV(1, 4) ROWV(1, 5) ROWV(2, 1) ROWV(2, 2) ROWV(2, 3)
    if (N <= 1024) { ROWV(4, 1) ROWV(4, 2) ROWV(4, 3) }
#define COLV(G, MB) { time_seq(c, row0, col_variant<K, G, MB>(), nrm0, us); report("col G=" #G " minb=" #MB, 1); }
    COLV(4, 1) COLV(4, 2) COLV(4, 3) COLV(4, 4) COLV(2, 2) COLV(2, 4) COLV(2, 6) COLV(8, 1) COLV(8, 2)
}

#endif

int main(int argc, char** argv) {
    Ctx c{};
    c.N = argc > 1 ? atoi(argv[1]) : 2048;
    c.count = argc > 2 ? atoi(argv[2]) : 1;
    c.reps = argc > 3 ? atoi(argv[3]) : 20;
    if (c.count > kMaxGroup) c.count = kMaxGroup;
    c.nsets = 8;
    while ((size_t)c.nsets * c.count * c.N * c.N * 48 > (size_t)4 << 30 && c.nsets > 1) c.nsets /= 2;
    const size_t nn = (size_t)c.N * c.N;
    float4* h0; float* ktab; CascadeDev* casc; float2* inter; float* disp; float4* normal; float* jac;
    CK(cudaMalloc(&h0, nn * sizeof(float4)));
    CK(cudaMalloc(&ktab, c.N * sizeof(float)));
    CK(cudaMalloc(&casc, sizeof(CascadeDev)));
    CK(cudaMalloc(&inter, nn / 2 * 3 * c.count * c.nsets * sizeof(float2)));
    CK(cudaMalloc(&disp, nn * 3 * c.count * c.nsets * sizeof(float)));
    CK(cudaMalloc(&normal, nn * c.count * c.nsets * sizeof(float4)));
    CK(cudaMalloc(&jac, nn * c.count * c.nsets * sizeof(float)));
    fill_h0<<<(unsigned)((nn + 255) / 256), 256>>>(h0, nn, 12345u);
    fill_ktab<<<(c.N + 255) / 256, 256>>>(ktab, c.N, 1000.0f);
    CascadeDev cd{1000.0f, 40.0f, 0.7071f, 0.7071f, 2.0f, 0.1f, 1.0f, 0.0f};
    CK(cudaMemcpy(casc, &cd, sizeof(cd), cudaMemcpyHostToDevice));
    float4 *hp, *nyq;
    CK(cudaMalloc(&hp, nn / 2 * 3 / 2 * sizeof(float4)));   // fold coefficients + (w, 1/|k|) table
    CK(cudaMalloc(&nyq, (size_t)(c.N / 2) * sizeof(float4)));
    fill_h0<<<(unsigned)((nn / 2 * 3 / 2 + 255) / 256), 256>>>(hp, nn / 2 * 3 / 2, 777u);     // timing only: any finite coefficients do
    fill_h0<<<(unsigned)((c.N / 2 + 255) / 256), 256>>>(nyq, c.N / 2, 778u);
    c.fb = FrameBuffers{c.N, h0, hp, nyq, ktab, casc, inter, disp, normal, jac, argc > 4 ? atoi(argv[4]) : 1};
    printf("discard_inter = %d\n", c.fb.discard_inter);
    for (int k = 0; k < c.nsets; ++k)
        for (int i = 0; i < c.count; ++i) { c.tab[k].cascade[i] = 0; c.tab[k].time[i] = 1.0f + i / 60.0f; c.tab[k].slot[i] = k * c.count + i; }
    for (auto& e : c.ev) CK(cudaEventCreate(&e));
    printf("== N=%d, %d frame(s) per launch, %d reps, %d rotating output sets (steady state of a sweep) ==\n", c.N, c.count, c.reps, c.nsets);
    switch (c.N) {
        case 512: sweep<512>(c); break;
        case 1024: sweep<1024>(c); break;
        case 2048: sweep<2048>(c); break;
        default: printf("N not instantiated in the tuner\n"); return 1;
    }
    return 0;
}
