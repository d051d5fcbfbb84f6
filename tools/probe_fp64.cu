// tools/probe_fp64.cu -- one-off B200 probe: the FP64 roofline denominators that MEASURED_PEAKS.json lacks.
// Measures: DFMA issue peak, DMMA (mma.sync f64) peaks for every shape, cuBLAS ZGEMM/DGEMM on the
// config-2 shapes (sanity ceiling only -- never on the product path), 16-byte copy bandwidth,
// launch and graph-node latencies.  Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuComplex.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m8n8k4: D(8x8) += A(8x4) B(4x8); per lane a:1 b:1 c:2 doubles
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters) {
    double c[ILP][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4: a:2 b:1 c:4 ; m16n8k8: a:4 b:2 c:4 ; m16n8k16: a:8 b:4 c:4
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1684(double *out, int iters) {
    double c[ILP][4];
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters) {
    double c[ILP][4];
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 * 2;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters) {
    double c[ILP][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-4 * i;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy16(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}
__global__ void k_empty(double *p) { if (p == nullptr) return; }
__global__ void k_tiny(double *p) { p[threadIdx.x] += 1.0; }

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"cc\":\"%d.%d\",\"clock_khz\":%d,\"l2_mb\":%.1f,\"smem_optin\":%zu}\n",
           p.name, p.multiProcessorCount, p.major, p.minor, clk, p.l2CacheSize / 1048576.0, p.sharedMemPerBlockOptin);
    const int nsm = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * 256 * nsm * 8));
    // ---- DFMA / DMMA peaks
    for (int occ : {1, 2, 4}) {
        const int grid = nsm * occ, iters = 20000;
        {
            float ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("{\"probe\":\"dfma\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", occ, fl / ms * 1e-9, ms);
        }
        {
            float ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters); }, 3);
            double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;
            printf("{\"probe\":\"dmma_m8n8k4\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", occ, fl / ms * 1e-9, ms);
        }
        {
            float ms = time_ms([&] { k_dmma1684<8><<<grid, 256>>>(out, iters); }, 3);
            double fl = 2.0 * 512 * 8 * iters * 8.0 * grid;
            printf("{\"probe\":\"dmma_m16n8k4\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", occ, fl / ms * 1e-9, ms);
        }
        {
            float ms = time_ms([&] { k_dmma1688<8><<<grid, 256>>>(out, iters); }, 3);
            double fl = 2.0 * 1024 * 8 * iters * 8.0 * grid;
            printf("{\"probe\":\"dmma_m16n8k8\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", occ, fl / ms * 1e-9, ms);
        }
        {
            float ms = time_ms([&] { k_dmma16816<8><<<grid, 256>>>(out, iters / 2); }, 3);
            double fl = 2.0 * 2048 * 8 * (iters / 2) * 8.0 * grid;
            printf("{\"probe\":\"dmma_m16n8k16\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", occ, fl / ms * 1e-9, ms);
        }
    }
    // ---- 16-byte copy bandwidth (4 GiB in, 4 GiB out: rank-14 tensor size)
    {
        size_t n = (size_t)1 << 28;
        double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 0, n * 16));
        float ms = time_ms([&] { k_copy16<<<nsm * 16, 512>>>(a, b, n); }, 5);
        printf("{\"probe\":\"copy16\",\"gbs\":%.1f,\"ms\":%.3f}\n", 2.0 * n * 16 / ms * 1e-6, ms);
        float ms2 = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); }, 5);
        printf("{\"probe\":\"memcpy_d2d\",\"gbs\":%.1f,\"ms\":%.3f}\n", 2.0 * n * 16 / ms2 * 1e-6, ms2);
        float ms3 = time_ms([&] { CK(cudaMemsetAsync(b, 0, n * 16)); }, 5);
        printf("{\"probe\":\"memset\",\"gbs\":%.1f,\"ms\":%.3f}\n", 1.0 * n * 16 / ms3 * 1e-6, ms3);
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    // ---- cuBLAS ZGEMM / DGEMM ceilings
    {
        cublasHandle_t h; cublasCreate(&h);
        struct Shape { int m, n, k; const char *tag; };
        std::vector<Shape> shapes = {{16384, 16384, 64, "cfg2_10x10"}, {4096, 65536, 64, "cfg2_9x11"},
                                     {64, 4194304, 64, "cfg2_6x14"}, {4194304, 64, 64, "cfg2_14x6"},
                                     {4096, 4096, 4096, "square4k"}, {8192, 8192, 8192, "square8k"},
                                     {16384, 16384, 16, "k16"}, {16384, 16384, 4, "k4"}};
        for (auto s : shapes) {
            cuDoubleComplex *A, *B, *C;
            CK(cudaMalloc(&A, (size_t)s.m * s.k * 16)); CK(cudaMalloc(&B, (size_t)s.k * s.n * 16)); CK(cudaMalloc(&C, (size_t)s.m * s.n * 16));
            CK(cudaMemset(A, 0, (size_t)s.m * s.k * 16)); CK(cudaMemset(B, 0, (size_t)s.k * s.n * 16));
            cuDoubleComplex one = make_cuDoubleComplex(1, 0), zero = make_cuDoubleComplex(0, 0);
            float ms = time_ms([&] { cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, s.m, s.n, s.k, &one, A, s.m, B, s.k, &zero, C, s.m); }, 3);
            double fl = 8.0 * s.m * s.n * s.k;
            double by = 16.0 * ((double)s.m * s.k + (double)s.k * s.n + (double)s.m * s.n);
            printf("{\"probe\":\"zgemm\",\"tag\":\"%s\",\"m\":%d,\"n\":%d,\"k\":%d,\"tflops\":%.2f,\"gbs\":%.1f,\"ms\":%.3f}\n",
                   s.tag, s.m, s.n, s.k, fl / ms * 1e-9, by / ms * 1e-6, ms);
            CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
        }
        {
            int n = 8192; double *A, *B, *C;
            CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
            CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
            double one = 1, zero = 0;
            float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 3);
            printf("{\"probe\":\"dgemm\",\"n\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", n, 2.0 * n * n * n / ms * 1e-9, ms);
            CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
        }
        cublasDestroy(h);
    }
    // ---- launch latency / graph node latency
    {
        const int n = 2000;
        float ms = time_ms([&] { for (int i = 0; i < n; i++) k_empty<<<1, 32>>>(out); }, 3);
        printf("{\"probe\":\"launch_empty\",\"us_per_launch\":%.3f}\n", ms * 1000 / n);
        ms = time_ms([&] { for (int i = 0; i < n; i++) k_tiny<<<1, 64>>>(out); }, 3);
        printf("{\"probe\":\"launch_tiny_dependent\",\"us_per_launch\":%.3f}\n", ms * 1000 / n);
        cudaStream_t st; CK(cudaStreamCreate(&st));
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < n; i++) k_tiny<<<1, 64, 0, st>>>(out);
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"probe\":\"graph_chain\",\"us_per_node\":%.3f}\n", ms * 1000 / n);
    }
    return 0;
}
