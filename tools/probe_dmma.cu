// tools/probe_dmma.cu -- what limits the complex DMMA inner loop of k_gett?  Variants of the fragment loop without
// global traffic: v2 real FXxFY tile with distinct operands, v3 complex 4-pass pattern (the k_gett loop, interleaved
// or pass-wise order), v4 = v3 + LDS.128 fragment loads from shared memory per k-step.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int FX, int FY>
__global__ void __launch_bounds__(512) v2(const double *in, double *out, int iters) {
    double a[FX], b[FY], c[FX][FY][2];
    for (int i = 0; i < FX; i++) a[i] = in[threadIdx.x + 32 * i];
    for (int j = 0; j < FY; j++) b[j] = in[threadIdx.x + 32 * j + 512];
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) c[i][j][0] = c[i][j][1] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    double s = 0;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int FX, int FY, bool LDS, int ORDER>
__global__ void __launch_bounds__(512) v3(const double2 *in, double *out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double2 xf[FX], yf[FY];
    double cr[FX][FY][2], ci[FX][FY][2];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < FX; i++) xf[i] = sm[t * 130 + i * 8 + g];
    for (int j = 0; j < FY; j++) yf[j] = sm[2100 + t * 66 + j * 8 + g];
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0;
    for (int it = 0; it < iters; it++) {
        if (LDS) {
            const int kk = it & 3;
#pragma unroll
            for (int i = 0; i < FX; i++) xf[i] = sm[(kk * 4 + t) * 130 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < FY; j++) yf[j] = sm[2100 + (kk * 4 + t) * 66 + j * 8 + g];
        }
        double nxi[FX];
#pragma unroll
        for (int i = 0; i < FX; i++) nxi[i] = -xf[i].y;
        if (ORDER == 0) {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) {
                    dmma(cr[i][j][0], cr[i][j][1], xf[i].x, yf[j].x);
                    dmma(ci[i][j][0], ci[i][j][1], xf[i].x, yf[j].y);
                    dmma(cr[i][j][0], cr[i][j][1], nxi[i], yf[j].y);
                    dmma(ci[i][j][0], ci[i][j][1], xf[i].y, yf[j].x);
                }
        } else {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(cr[i][j][0], cr[i][j][1], xf[i].x, yf[j].x);
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(ci[i][j][0], ci[i][j][1], xf[i].x, yf[j].y);
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(cr[i][j][0], cr[i][j][1], nxi[i], yf[j].y);
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(ci[i][j][0], ci[i][j][1], xf[i].y, yf[j].x);
        }
    }
    double s = 0;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) s += cr[i][j][0] + cr[i][j][1] + ci[i][j][0] + ci[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// v5: the 3M loop of k_gett (T1 = Xr*Yr, T2 = Xi*Yi, T3 = (Xr+Xi)*(Yr+Yi)): LDS fragments + DADD sums + 3 DMMA passes.
// SUMS = 0 drops the DADDs (wrong maths, same DMMA count) to isolate what the FP64 adds cost next to the DMMAs.
template <int FX, int FY, int SUMS>
__global__ void __launch_bounds__(512) v5(const double2 *in, double *out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double c1[FX][FY][2], c2[FX][FY][2], c3[FX][FY][2];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = c3[i][j][0] = c3[i][j][1] = 0;
    for (int it = 0; it < iters; it++) {
        const int kk = it & 3;
        double2 xf[FX], yf[FY];
        double xs[FX], ys[FY];
#pragma unroll
        for (int i = 0; i < FX; i++) xf[i] = sm[(kk * 4 + t) * 66 + i * 8 + g];
#pragma unroll
        for (int j = 0; j < FY; j++) yf[j] = sm[2100 + (kk * 4 + t) * 66 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < FX; i++) xs[i] = SUMS ? xf[i].x + xf[i].y : xf[i].x;
#pragma unroll
        for (int j = 0; j < FY; j++) ys[j] = SUMS ? yf[j].x + yf[j].y : yf[j].y;
#pragma unroll
        for (int i = 0; i < FX; i++)
#pragma unroll
            for (int j = 0; j < FY; j++) dmma(c1[i][j][0], c1[i][j][1], xf[i].x, yf[j].x);
#pragma unroll
        for (int i = 0; i < FX; i++)
#pragma unroll
            for (int j = 0; j < FY; j++) dmma(c2[i][j][0], c2[i][j][1], xf[i].y, yf[j].y);
#pragma unroll
        for (int i = 0; i < FX; i++)
#pragma unroll
            for (int j = 0; j < FY; j++) dmma(c3[i][j][0], c3[i][j][1], xs[i], ys[j]);
    }
    double s = 0;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) s += c1[i][j][0] + c1[i][j][1] + c2[i][j][0] + c2[i][j][1] + c3[i][j][0] + c3[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// v6: v5 with the re+im sums issued back to back (asm volatile keeps program order), GROUP = how many k-steps' worth of
// sums are computed in one burst (1: 4 DADDs then 12 DMMAs; 4: 16 DADDs then 48 DMMAs) -- is the ~4.7-cycle cost of a DADD
// next to DMMAs a pipe turn-around that batching amortises?
__device__ __forceinline__ double dadd_v(double a, double b) { double r; asm volatile("add.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r; }
template <int GROUP>
__global__ void __launch_bounds__(512) v6(const double2 *in, double *out, int iters) {
    constexpr int FX = 2, FY = 2;
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double c1[FX][FY][2], c2[FX][FY][2], c3[FX][FY][2];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) c1[i][j][0] = c1[i][j][1] = c2[i][j][0] = c2[i][j][1] = c3[i][j][0] = c3[i][j][1] = 0;
    for (int it = 0; it < iters; it += GROUP) {
        double2 xf[GROUP][FX], yf[GROUP][FY];
        double xs[GROUP][FX], ys[GROUP][FY];
#pragma unroll
        for (int u = 0; u < GROUP; u++) {
            const int kk = (it + u) & 3;
#pragma unroll
            for (int i = 0; i < FX; i++) xf[u][i] = sm[(kk * 4 + t) * 66 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < FY; j++) yf[u][j] = sm[2100 + (kk * 4 + t) * 66 + j * 8 + g];
        }
#pragma unroll
        for (int u = 0; u < GROUP; u++) {
#pragma unroll
            for (int i = 0; i < FX; i++) xs[u][i] = dadd_v(xf[u][i].x, xf[u][i].y);
#pragma unroll
            for (int j = 0; j < FY; j++) ys[u][j] = dadd_v(yf[u][j].x, yf[u][j].y);
        }
        __syncwarp();                                   // scheduling fence: keeps the DADD burst together in SASS
#pragma unroll
        for (int u = 0; u < GROUP; u++) {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(c1[i][j][0], c1[i][j][1], xf[u][i].x, yf[u][j].x);
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(c2[i][j][0], c2[i][j][1], xf[u][i].y, yf[u][j].y);
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) dmma(c3[i][j][0], c3[i][j][1], xs[u][i], ys[u][j]);
        }
    }
    double s = 0;
    for (int i = 0; i < FX; i++) for (int j = 0; j < FY; j++) s += c1[i][j][0] + c1[i][j][1] + c2[i][j][0] + c2[i][j][1] + c3[i][j][0] + c3[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// v7: how fast is a stream of DADDs alone / DADD:DMMA mixes from ONE warp per scheduler vs four (R DADDs per DMMA)
template <int R>
__global__ void __launch_bounds__(512) v7(const double *in, double *out, int iters) {
    double a = in[threadIdx.x], b = in[threadIdx.x + 512];
    double c0 = 0, c1 = 0, e[8];
    for (int i = 0; i < 8; i++) e[i] = in[threadIdx.x + 32 * i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            dmma(c0, c1, a, b);
#pragma unroll
            for (int q = 0; q < R; q++) e[(r * R + q) & 7] = dadd_v(e[(r * R + q) & 7], a);
        }
    }
    double s = c0 + c1;
    for (int i = 0; i < 8; i++) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

int main() {
    double *in, *out; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, 1 << 24)); CK(cudaMemset(in, 0, 1 << 20));
    const int nsm = 148, iters = 4000;
    const size_t smem = 4096 * 16;
    auto rep = [&](const char *name, int warps, int dm, float ms) {
        double fl = 2.0 * 256 * dm * (double)iters * warps * nsm;
        printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", name, warps, fl / ms * 1e-9, ms);
    };
#define RUN3(FX, FY, L, O, NAME) \
    CK(cudaFuncSetAttribute(v3<FX, FY, L, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int w : {8, 16}) { float ms = time_ms([&] { v3<FX, FY, L, O><<<nsm, w * 32, smem>>>((const double2 *)in, out, iters); }); rep(NAME, w, 4 * FX * FY, ms); }
    for (int w : {8, 16}) { float ms = time_ms([&] { v2<4, 4><<<nsm, w * 32>>>(in, out, iters); }); rep("v2_real_4x4_distinct_operands", w, 64, ms); }
    for (int w : {8, 16}) { float ms = time_ms([&] { v2<4, 2><<<nsm, w * 32>>>(in, out, iters); }); rep("v2_real_4x2", w, 32, ms); }
    RUN3(4, 4, false, 0, "v3_complex_4x4_interleaved_noLDS");
    RUN3(4, 4, false, 1, "v3_complex_4x4_4pass_noLDS");
    RUN3(4, 4, true, 1, "v4_complex_4x4_4pass_LDS");
    RUN3(4, 2, false, 1, "v3_complex_4x2_4pass_noLDS");
    RUN3(4, 2, true, 1, "v4_complex_4x2_4pass_LDS");
    RUN3(2, 2, true, 1, "v4_complex_2x2_4pass_LDS");
#define RUN5(FX, FY, S, NAME) \
    CK(cudaFuncSetAttribute(v5<FX, FY, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int w : {8, 16}) { float ms = time_ms([&] { v5<FX, FY, S><<<nsm, w * 32, smem>>>((const double2 *)in, out, iters); }); rep(NAME, w, 3 * FX * FY, ms); }
    RUN5(2, 2, 1, "v5_3M_2x2_LDS_DADD");
    RUN5(2, 2, 0, "v5_3M_2x2_LDS_noDADD");
    RUN5(4, 2, 1, "v5_3M_4x2_LDS_DADD");
    RUN5(4, 2, 0, "v5_3M_4x2_LDS_noDADD");
#define RUN6(G, NAME) \
    CK(cudaFuncSetAttribute(v6<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int w : {8, 16}) { float ms = time_ms([&] { v6<G><<<nsm, w * 32, smem>>>((const double2 *)in, out, iters); }); rep(NAME, w, 12, ms); }
    RUN6(1, "v6_3M_2x2_sums_burst_of_4");
    RUN6(2, "v6_3M_2x2_sums_burst_of_8");
    RUN6(4, "v6_3M_2x2_sums_burst_of_16");
    // v7: 4 DMMAs + 4R DADDs per iteration; report the time per iteration in SM cycles per scheduler-warp
    auto rep7 = [&](const char *name, int warps, int R, float ms) {
        double cyc = ms * 1e-3 * 1.965e9 / iters;          // cycles per iteration (4 DMMA + 4R DADD per warp)
        printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"dadd_per_dmma\":%d,\"cycles_per_iter\":%.1f,\"ms\":%.3f}\n", name, warps, R, cyc, ms);
    };
    for (int w : {4, 16}) { float ms = time_ms([&] { v7<0><<<nsm, w * 32>>>(in, out, iters); }); rep7("v7_dmma_only", w, 0, ms); }
    for (int w : {4, 16}) { float ms = time_ms([&] { v7<1><<<nsm, w * 32>>>(in, out, iters); }); rep7("v7_dmma_dadd", w, 1, ms); }
    for (int w : {4, 16}) { float ms = time_ms([&] { v7<2><<<nsm, w * 32>>>(in, out, iters); }); rep7("v7_dmma_dadd", w, 2, ms); }
    for (int w : {4, 16}) { float ms = time_ms([&] { v7<4><<<nsm, w * 32>>>(in, out, iters); }); rep7("v7_dmma_dadd", w, 4, ms); }
    return 0;
}
