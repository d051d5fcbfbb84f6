// qtorch_b200/csrc/reduce.cuh -- split-K kernel for steps with (almost) no free legs and a long sum,
// e.g. the (14,14,k=14 -> 0) inner product that closes the QAOA-30 line-graph plan
// (8.6 GB read once: HBM-bound).  Replaces the same loop of Network::ContractIndices
// (/root/reference/src/Network.h:892-935) for rC <= 2.
//
//   C[c] = sum_s A[fA(c) + sA(s)] * B[fB(c) + sB(s)]
// The summed index is cut into tiles of 256 (four legs chosen so that BOTH operands are read in
// >= 256-byte runs); CTAs stride over tiles, every thread keeps NC complex accumulators, a block
// reduction writes one partial per CTA and a second tiny kernel adds the partials in a fixed order
// (deterministic, no atomics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace qtb {

struct ReduceParams {
    const double2 *A;
    const double2 *B;
    double2 *partial;          // [gridDim.x][NC]
    double2 *C;
    uint32_t nTiles;           // 4^k / 256
    uint8_t kbits;             // 2k
    uint8_t shA[32], shB[32];  // summed bit j -> bit position inside A / B (bits 0..7 = tile)
    uint32_t fA[16], fB[16];   // offsets of output element c inside A / B
};

__device__ __forceinline__ uint32_t scatter32(uint32_t v, const uint8_t *sh, int first, int count) {
    uint32_t o = 0;
    for (int j = 0; j < count; j++) o += ((v >> j) & 1u) << sh[first + j];
    return o;
}

template <int NC>
__global__ void __launch_bounds__(256) k_reduce(const ReduceParams p) {
    __shared__ double2 red[8][NC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tA = scatter32(tid, p.shA, 0, 8), tB = scatter32(tid, p.shB, 0, 8);
    const int hi = p.kbits - 8;
    double accR[NC], accI[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { accR[c] = 0.0; accI[c] = 0.0; }
    const double2 *__restrict__ A = p.A;
    const double2 *__restrict__ B = p.B;
#pragma unroll 2
    for (uint32_t tile = blockIdx.x; tile < p.nTiles; tile += gridDim.x) {
        const uint32_t oa = scatter32(tile, p.shA, 8, hi) + tA, ob = scatter32(tile, p.shB, 8, hi) + tB;
#pragma unroll
        for (int c = 0; c < NC; c++) cmac(accR[c], accI[c], A[(size_t)p.fA[c] + oa], B[(size_t)p.fB[c] + ob]);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const double r = warp_sum(accR[c]), i = warp_sum(accI[c]);
        if (lane == 0) red[warp][c] = make_double2(r, i);
    }
    __syncthreads();
    if (tid < NC) {
        double r = 0.0, i = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) { r += red[w][tid].x; i += red[w][tid].y; }
        p.partial[(size_t)blockIdx.x * NC + tid] = make_double2(r, i);
    }
}

// ---------------------------------------------------------------------------------------------
// rC = 0 (pure inner product with a digit-permuted second operand): both operands are read in their
// OWN memory order (lanes follow each operand's contiguous bits -> 256-byte runs for A and for B) and
// meet through shared memory, where the tile coordinate undoes the permutation.
struct DotParams {
    const double2 *A;
    const double2 *B;
    double2 *partial;
    uint32_t nTiles;
    uint8_t kbits;
    uint8_t shA[32], shB[32];      // summed bit j (tile-coordinate order) -> bit position inside A / B
    uint8_t permA[8], permB[8];    // thread bit j -> tile-coordinate bit, ordered by the operand's memory significance
};

__device__ __forceinline__ uint32_t dot_swz(uint32_t c) { return c ^ ((c >> 3) & 7u) ^ ((c >> 6) & 3u); }

#define QTB_DOT_T 4
__global__ void __launch_bounds__(256) k_dot(const DotParams p) {
    __shared__ double2 sA[QTB_DOT_T][256], sB[QTB_DOT_T][256];
    __shared__ double2 red[8];
    __shared__ uint32_t hA[4][64], hB[4][64];      // tile-index bits -> offsets, 6 bits per table (k <= 16: 24 bits)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t cA = 0, cB = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { cA |= ((tid >> j) & 1u) << p.permA[j]; cB |= ((tid >> j) & 1u) << p.permB[j]; }
    const uint32_t oA = scatter32(cA, p.shA, 0, 8), oB = scatter32(cB, p.shB, 0, 8);
    const uint32_t wA = dot_swz(cA), wB = dot_swz(cB), rd = dot_swz(tid);
    const int hi = p.kbits - 8;
    {
        const int part = tid >> 6, v = tid & 63, lo = 6 * part;
        const int n = hi - lo < 0 ? 0 : (hi - lo > 6 ? 6 : hi - lo);
        hA[part][v] = scatter32(v, p.shA, 8 + lo, n);
        hB[part][v] = scatter32(v, p.shB, 8 + lo, n);
    }
    __syncthreads();
    const double2 *__restrict__ A = p.A;
    const double2 *__restrict__ B = p.B;
    double accR = 0.0, accI = 0.0;
    for (uint32_t base = blockIdx.x * QTB_DOT_T; base < p.nTiles; base += gridDim.x * QTB_DOT_T) {
        double2 a[QTB_DOT_T], b[QTB_DOT_T];
#pragma unroll
        for (int u = 0; u < QTB_DOT_T; u++) {
            const uint32_t tile = base + u;
            if (tile < p.nTiles) {
                const uint32_t t0 = tile & 63, t1 = (tile >> 6) & 63, t2 = (tile >> 12) & 63, t3 = (tile >> 18) & 63;
                a[u] = A[(size_t)(hA[0][t0] + hA[1][t1] + hA[2][t2] + hA[3][t3]) + oA];
                b[u] = B[(size_t)(hB[0][t0] + hB[1][t1] + hB[2][t2] + hB[3][t3]) + oB];
            } else {
                a[u] = make_double2(0.0, 0.0); b[u] = make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < QTB_DOT_T; u++) { sA[u][wA] = a[u]; sB[u][wB] = b[u]; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QTB_DOT_T; u++) cmac(accR, accI, sA[u][rd], sB[u][rd]);
        __syncthreads();
    }
    const double r = warp_sum(accR), i = warp_sum(accI);
    if (lane == 0) red[warp] = make_double2(r, i);
    __syncthreads();
    if (tid == 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) { sr += red[w].x; si += red[w].y; }
        p.partial[blockIdx.x] = make_double2(sr, si);
    }
}

template <int NC>
__global__ void __launch_bounds__(32 * NC) k_reduce_final(const double2 *__restrict__ partial, double2 *C, uint32_t nBlocks) {
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    double r = 0.0, i = 0.0;
    for (uint32_t b = lane; b < nBlocks; b += 32) { const double2 v = partial[(size_t)b * NC + c]; r += v.x; i += v.y; }
    r = warp_sum(r); i = warp_sum(i);
    if (lane == 0) C[c] = make_double2(r, i);
}

}  // namespace qtb
