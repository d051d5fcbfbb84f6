// qtorch_b200/csrc/apply.cuh -- streaming steps: a big Node tensor contracted with a tiny one (<= 64 elements).
//
// The "gate / measurement cap applied to a large intermediate" shape of Network::ContractIndices
// (/root/reference/src/Network.h:892-935 with rB <= 3): arithmetic intensity <= 2 flop/B, so the step is a pure
// HBM stream -- read the big operand once, write the result once.  One thread per free index x of the big operand:
//      C[x, y] = sum_s X[ins(x) + kOff[s]] * W[s][y]            K = 4^k <= 16 loads of 16 B, N = 4^nfy <= 16 outputs
// ins(x) opens the (at most two) 2-bit holes of the shared legs in x, W (the small operand re-ordered as [s][y]) sits in
// shared memory and is read as a broadcast.  Consecutive threads read consecutive elements of X for every s and write
// consecutive elements of C for every y; when the small operand's legs come first in C (C index = y + N x) the block's
// 256 N outputs form one contiguous run and are transposed through shared memory so that the stores stream as well.
// No tile staging, no tensor pipe (the DMMA tile kernel pads N to 8 and idles on these).  N <= 16, K <= 16.
// When the small operand's legs come first in C a thread's N outputs are ONE contiguous run of 16 N bytes: it is written
// with 256-bit stores (whole 32-byte sectors per instruction), no shared-memory transpose.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qtb {

struct ApplyParams {
    const double2 *X;             // big operand, original layout
    const double2 *Y;             // small operand (<= 64 elements)
    double2 *C;
    uint64_t M;                   // free elements of X = number of threads
    uint8_t nHoles;               // = k (0, 1 or 2)
    uint8_t holeBit[2];           // bit position of each shared leg inside X's element index, ascending
    uint8_t yFirst;               // 1: C index = y + N * x (small operand is the reference's node A), 0: x + M * y
    uint8_t lowRun;               // K == 4 and the shared leg is X's leg 0: the four summands are one contiguous 64-byte run
    uint32_t kOff[16];            // element offset inside X of summed value s
    uint8_t yIdx[16][16];         // element index inside Y of (s, y)
};

__device__ __forceinline__ void st_stream_256(double2 *dst, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ void ld_stream_256(const double2 *src, double2 &a, double2 &b) {
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(src) : "memory");
}

template <int K, int N>
__global__ void __launch_bounds__(256) k_apply(const ApplyParams p) {
    __shared__ double2 W[K * N];
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) W[i] = p.Y[p.yIdx[i / N][i % N]];
    __syncthreads();
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;        // M is a multiple of the block size
    uint64_t off = x;
#pragma unroll
    for (int j = 0; j < 2; j++)
        if (j < p.nHoles) { const int h = p.holeBit[j]; off = ((off >> h) << (h + 2)) | (off & ((1ull << h) - 1)); }
    double2 xv[K];
    if (K == 4 && p.lowRun) {
        // kOff = {0, 1, 2, 3}: whole sectors with two 256-bit loads instead of four 16-byte loads a sector apart
        ld_stream_256(p.X + off, xv[0], xv[1]);
        ld_stream_256(p.X + off + 2, xv[2 % K], xv[3 % K]);
    } else {
#pragma unroll
        for (int s = 0; s < K; s++) xv[s] = __ldcs(p.X + off + p.kOff[s]);        // streamed once: evict-first
    }
    double accR[N], accI[N];
#pragma unroll
    for (int y = 0; y < N; y++) accR[y] = accI[y] = 0.0;
#pragma unroll
    for (int s = 0; s < K; s++)
#pragma unroll
        for (int y = 0; y < N; y++) {
            const double2 w = W[s * N + y];
            accR[y] = fma(xv[s].x, w.x, accR[y]);
            accR[y] = fma(-xv[s].y, w.y, accR[y]);
            accI[y] = fma(xv[s].x, w.y, accI[y]);
            accI[y] = fma(xv[s].y, w.x, accI[y]);
        }
    if (N > 1 && p.yFirst) {
        double2 *c = p.C + x * N;                                              // this thread's run: y fastest
#pragma unroll
        for (int y = 0; y < N; y += 2) st_stream_256(c + y, accR[y], accI[y], accR[y + 1], accI[y + 1]);
    } else {
#pragma unroll
        for (int y = 0; y < N; y++) __stcs(p.C + x + p.M * y, make_double2(accR[y], accI[y]));
    }
}

// Both shared legs are the big operand's two lowest legs: the 16 summands of an output are ONE 256-byte run.  Four lanes
// share an output -- lane q reads the 64-byte quarter q of the run (so a warp's requests tile 2 KB without gaps), sums
// its four terms, and two shuffle steps add the quarters.  yIdx / W are indexed by MEMORY offset m = 4 q + i here.
template <int N>
__global__ void __launch_bounds__(256) k_apply_lowpair(const ApplyParams p) {
    // W[q][i][y] with one element of padding per quarter: the four quarters a warp reads side by side sit in different banks
    constexpr int QS = 4 * N + 1;
    __shared__ double2 W[4 * QS];
    for (int i = threadIdx.x; i < 16 * N; i += blockDim.x) W[(i / (4 * N)) * QS + i % (4 * N)] = p.Y[p.yIdx[i / N][i % N]];
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;       // 4 M threads, a multiple of the block size
    const uint64_t x = t >> 2;
    const int q = (int)(t & 3);
    const double2 *src = p.X + (x << 4) + 4 * q;
    double2 xv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xv[i] = __ldcs(src + i);
    double accR[N], accI[N];
#pragma unroll
    for (int y = 0; y < N; y++) accR[y] = accI[y] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int y = 0; y < N; y++) {
            const double2 w = W[q * QS + i * N + y];
            accR[y] = fma(xv[i].x, w.x, accR[y]);
            accR[y] = fma(-xv[i].y, w.y, accR[y]);
            accI[y] = fma(xv[i].x, w.y, accI[y]);
            accI[y] = fma(xv[i].y, w.x, accI[y]);
        }
#pragma unroll
    for (int y = 0; y < N; y++) {
        accR[y] += __shfl_xor_sync(0xffffffffu, accR[y], 1); accI[y] += __shfl_xor_sync(0xffffffffu, accI[y], 1);
        accR[y] += __shfl_xor_sync(0xffffffffu, accR[y], 2); accI[y] += __shfl_xor_sync(0xffffffffu, accI[y], 2);
    }
    if (q == 0) {
#pragma unroll
        for (int y = 0; y < N; y++) __stcs(p.yFirst ? p.C + x * N + y : p.C + x + p.M * y, make_double2(accR[y], accI[y]));
    }
}

}  // namespace qtb
