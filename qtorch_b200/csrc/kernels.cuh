// qtorch_b200/csrc/kernels.cuh -- sm_100a kernels for the latency- and bandwidth-bound step classes.
//
//   k_step_thread  one thread per output element, sequential sum over the K = 4^k shared index
//                  (mid-size steps; coalesced along C / A's low free legs)
//   k_step_warp    one warp per output element, lanes stride over K, shuffle reduction
//                  (few outputs, long sums: the rank-0/1 results that close a network)
//   k_micro        the grouped micro-step executor: ONE launch runs a whole dependency-levelled
//                  list of tiny steps (replaces thousands of ContractNodes-sized launches; the
//                  GHZ-1000 plan is 2 999 steps of <= 4^5 MACs, /root/reference/src/Network.h:876)
//
// All of them compute  C[c] = sum_s A[offA(c,s)] * B[offB(c,s)]  (Network.h:892-935) with FP64 FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "step.h"

namespace qtb {

__device__ __forceinline__ void cmac(double &cr, double &ci, const double2 a, const double2 b) {
    cr = fma(a.x, b.x, cr);
    cr = fma(-a.y, b.y, cr);
    ci = fma(a.x, b.y, ci);
    ci = fma(a.y, b.x, ci);
}

// element offsets contributed by output index c
__device__ __forceinline__ void free_offsets(const DevStep &st, uint64_t c, uint64_t &oa, uint64_t &ob) {
    oa = 0; ob = 0;
    const int nfa = st.nfa, rC = st.rC;
    for (int i = 0; i < nfa; i++) oa += ((c >> (2 * i)) & 3ull) << st.shFree[i];
    for (int i = nfa; i < rC; i++) ob += ((c >> (2 * i)) & 3ull) << st.shFree[i];
}
// element offsets contributed by summed index s
__device__ __forceinline__ void sum_offsets(const DevStep &st, uint64_t s, uint64_t &oa, uint64_t &ob) {
    oa = 0; ob = 0;
    const int k = st.k;
    for (int i = 0; i < k; i++) {
        const uint64_t d = (s >> (2 * i)) & 3ull;
        oa += d << st.shSumA[i];
        ob += d << st.shSumB[i];
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_step_thread(const DevStep st) {
    const uint64_t NC = 1ull << (2 * st.rC), K = 1ull << (2 * st.k);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < NC; c += stride) {
        uint64_t ba, bb;
        free_offsets(st, c, ba, bb);
        const double2 *__restrict__ pa = st.A + ba;
        const double2 *__restrict__ pb = st.B + bb;
        double cr = 0.0, ci = 0.0;
        if (st.k <= 3) {
            // small K: offsets of the <= 64 summed terms are cheap to rebuild digit by digit
            for (uint64_t s = 0; s < K; s++) {
                uint64_t oa, ob;
                sum_offsets(st, s, oa, ob);
                cmac(cr, ci, pa[oa], pb[ob]);
            }
        } else {
            // long sums: walk the low three summed digits in an inner block of 64
            for (uint64_t s1 = 0; s1 < K; s1 += 64) {
                uint64_t oa1, ob1;
                sum_offsets(st, s1, oa1, ob1);
#pragma unroll 4
                for (uint32_t s0 = 0; s0 < 64; s0++) {
                    const uint64_t oa = oa1 + ((uint64_t)(s0 & 3) << st.shSumA[0]) + ((uint64_t)((s0 >> 2) & 3) << st.shSumA[1]) +
                                        ((uint64_t)(s0 >> 4) << st.shSumA[2]);
                    const uint64_t ob = ob1 + ((uint64_t)(s0 & 3) << st.shSumB[0]) + ((uint64_t)((s0 >> 2) & 3) << st.shSumB[1]) +
                                        ((uint64_t)(s0 >> 4) << st.shSumB[2]);
                    cmac(cr, ci, pa[oa], pb[ob]);
                }
            }
        }
        st.C[c] = make_double2(cr, ci);
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) k_step_warp(const DevStep st) {
    const uint64_t NC = 1ull << (2 * st.rC), K = 1ull << (2 * st.k);
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t c = warp; c < NC; c += nwarps) {
        uint64_t ba, bb;
        free_offsets(st, c, ba, bb);
        double cr = 0.0, ci = 0.0;
        for (uint64_t s = lane; s < K; s += 32) {
            uint64_t oa, ob;
            sum_offsets(st, s, oa, ob);
            cmac(cr, ci, st.A[ba + oa], st.B[bb + ob]);
        }
        cr = warp_sum(cr); ci = warp_sum(ci);
        if (lane == 0) st.C[c] = make_double2(cr, ci);
    }
}

// ---------------------------------------------------------------------------------------------
// Grouped micro-step executor.
//
// blob layout (built by the host, one per independent plan; blockIdx.x selects the blob):
//   MicroHeader | uint32 levelItemStart[nLevels+1] | MicroItem items[nItems] | DevStep steps[nSteps]
// Steps of one dependency level are independent; a level's work is cut into items of
// QTB_MICRO_CHUNK outputs which the CTA's warps share; __syncthreads() separates levels (it also
// orders the global-memory writes of one level before the reads of the next, CTA scope).
#define QTB_MICRO_CHUNK 128
#define QTB_MICRO_THREADS 1024         // single plans (latency-bound chains of tiny steps: GHZ-1000 is 2 999 levels of one item)
#define QTB_MICRO_THREADS_BATCH 512    // term batches: 16 warps x 128 registers, a lane keeps 16 operand loads in flight; measured with clusters: p=2 objective 0.257 ms (1024) -> 0.198 ms (512)

struct MicroHeader {
    uint32_t nLevels, nItems, nSteps, stepsOffset;
    uint32_t controlBytes;        // header + level table + items + steps (what the kernel stages in shared memory)
    uint32_t prefetchBytes;       // size of the operand region to pull into L2 up front (0 = none)
    uint64_t prefetchPtr;         // device address of that region (plan input blob / upload payload)
    uint64_t timelinePtr;         // debugging (QTB_MICRO_TIMELINE=1): device array of nLevels + 2 clock64() stamps, else 0
};
struct MicroItem { uint32_t step; uint32_t chunk; };      // chunk: bits 0-23 index of the item inside its step, bits 24-28 log2 of the lanes per output, bits 29-30 log2 of the passes

// per-warp scratch in dynamic shared memory: the step descriptor (so leg tables are not re-read from global memory
// for every index) and the offsets of all summed terms when K <= QTB_MICRO_TAB
#define QTB_MICRO_TAB 256
struct MicroWarpScratch {
    DevStep st;
    uint32_t sumA[QTB_MICRO_TAB], sumB[QTB_MICRO_TAB];
};
static_assert(sizeof(DevStep) % 4 == 0, "DevStep is copied word-wise");
#define QTB_MICRO_CTRL_BYTES (96 * 1024)       // control structures up to this size are staged in shared memory
#define QTB_MICRO_SMEM_FOR(T) (((T) / 32) * sizeof(MicroWarpScratch) + QTB_MICRO_CTRL_BYTES)
#define QTB_MICRO_SMEM QTB_MICRO_SMEM_FOR(QTB_MICRO_THREADS)

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_micro_t(const uint8_t *__restrict__ blobBase,
                                                                  const uint64_t *__restrict__ blobOffsets) {
    extern __shared__ __align__(16) uint8_t microSmem[];
    // A plan may be run by a thread-block CLUSTER (launch attribute; 1 x 1 x 1 when launched plainly): the CTAs of a cluster share
    // the items of every level -- a level of a QAOA p=2 term is bound by one SM's load/store unit (scattered 16-byte operand reads:
    // 780 cycles per multiply-add and warp, tools/micro_timeline.py), which more SMs multiply -- and meet in a cluster barrier
    // between levels (release / acquire at cluster scope orders the global-memory writes of a level before the next level's reads).
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const uint8_t *blob = blobBase + blobOffsets[blockIdx.x / csize];
    const MicroHeader hdr = *reinterpret_cast<const MicroHeader *>(blob);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    MicroWarpScratch &ws = reinterpret_cast<MicroWarpScratch *>(microSmem)[warp];
    // A launch usually starts cold (the big steps in between stream gigabytes through L2), and every item would
    // otherwise chase item -> descriptor -> operand through DRAM one round trip at a time: stage the control
    // structures in shared memory with one cooperative read and pull the operand region into L2 meanwhile.
    uint8_t *ctrl = microSmem + (THREADS / 32) * sizeof(MicroWarpScratch);
    const bool staged = hdr.controlBytes <= QTB_MICRO_CTRL_BYTES;
    if (staged) {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(ctrl);
        for (uint32_t i = threadIdx.x; i < (hdr.controlBytes + 15) / 16; i += blockDim.x) dst[i] = src[i];
    }
    if (hdr.prefetchBytes) {
        const uint8_t *pf = reinterpret_cast<const uint8_t *>(hdr.prefetchPtr);
        for (uint32_t off = threadIdx.x * 128u; off < hdr.prefetchBytes; off += blockDim.x * 128u)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + off));
    }
    __syncthreads();
    unsigned long long *timeline = reinterpret_cast<unsigned long long *>(hdr.timelinePtr);
    if (timeline && threadIdx.x == 0 && crank == 0) timeline[0] = clock64();
    const uint8_t *cb = staged ? ctrl : blob;
    const uint32_t *lis = reinterpret_cast<const uint32_t *>(cb + sizeof(MicroHeader));
    const MicroItem *items = reinterpret_cast<const MicroItem *>(lis + hdr.nLevels + 1);
    const DevStep *steps = reinterpret_cast<const DevStep *>(cb + hdr.stepsOffset);

    for (uint32_t lvl = 0; lvl < hdr.nLevels; lvl++) {
        const uint32_t i0 = lis[lvl], i1 = lis[lvl + 1];
        for (uint32_t it = i0 + crank * nw + warp; it < i1; it += nw * csize) {
            const MicroItem item = items[it];
            // descriptor -> this warp's shared-memory slot (one coalesced read)
            __syncwarp();
            {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&steps[item.step]);
                uint32_t *dst = reinterpret_cast<uint32_t *>(&ws.st);
                for (int w = lane; w < (int)(sizeof(DevStep) / 4); w += 32) dst[w] = src[w];
            }
            __syncwarp();
            const DevStep &st = ws.st;
            const uint32_t NC = 1u << (2 * st.rC), K = 1u << (2 * st.k);
            // plain (coherent) loads: operands may have been written earlier in this launch
            const double2 *A = st.A;
            const double2 *B = st.B;
            if (st.kind == KIND_COPY) {
                // staged upload: payload inside this blob -> the tensor's pooled buffer
                // (element count travels in the B slot)
                const uint32_t elems = (uint32_t)reinterpret_cast<uintptr_t>(B);
#pragma unroll
                for (int j = 0; j < QTB_MICRO_CHUNK / 32; j++) {
                    const uint32_t e = item.chunk * QTB_MICRO_CHUNK + j * 32 + lane;
                    if (e < elems) st.C[e] = A[e];
                }
                continue;
            }
            // offsets of the summed terms: a table for the low four summed digits (all of them when K <= 256); a longer sum walks
            // it once per block of 256 terms, adding the block's own offset (digit contributions are additive)
            const uint32_t KT = K < QTB_MICRO_TAB ? K : QTB_MICRO_TAB;
            for (uint32_t s = lane; s < KT; s += 32) {
                uint64_t oa, ob;
                sum_offsets(st, s, oa, ob);
                ws.sumA[s] = (uint32_t)oa; ws.sumB[s] = (uint32_t)ob;
            }
            __syncwarp();
            // lane layout (chosen by the host, build_micro_blob): G = 2^lg lanes cover one output along the summed index and a
            // shuffle tree adds their slices; P = 32 / G outputs per pass, 2^lp <= 4 passes per item.  Tiny results (NC < 32) and
            // steps with few outputs but long sums use the whole warp -- and more warps of the CTA -- this way.  The summed loop
            // is unrolled so several independent loads are in flight.
            const uint32_t lg = (item.chunk >> 24) & 31u, lp = item.chunk >> 29, chunk = item.chunk & 0xffffffu;
            const uint32_t G = 1u << lg, P = 32u >> lg;
            const uint32_t sg = (uint32_t)lane >> (5 - lg);
            const int passes = 1 << lp;
#pragma unroll 1
            for (int j = 0; j < passes; j++) {
                const uint32_t c = ((chunk << lp) + (uint32_t)j) * P + ((uint32_t)lane & (P - 1));
                double cr = 0.0, ci = 0.0;
                if (c < NC) {
                    uint64_t ba, bb;
                    free_offsets(st, c, ba, bb);
                    const double2 *pa = A + ba, *pb = B + bb;
                    // a step's time is its serial chain of memory round trips (tools/micro_timeline.py: 450-800 cycles per
                    // multiply-add when ptxas re-used the load registers, 770 when every term's offsets were rebuilt digit by
                    // digit): eight terms per round -- sixteen loads in flight -- then four, then the rest
                    for (uint32_t hi = 0; hi < K; hi += QTB_MICRO_TAB) {
                        const double2 *pah = pa, *pbh = pb;
                        if (K > QTB_MICRO_TAB) {
                            uint64_t oaH, obH;
                            sum_offsets(st, hi, oaH, obH);
                            pah += oaH; pbh += obH;
                        }
                        uint32_t s = sg;
                        for (; s + 7 * G < KT; s += 8 * G) {
                            double2 va[8], vb[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) { va[u] = pah[ws.sumA[s + u * G]]; vb[u] = pbh[ws.sumB[s + u * G]]; }
#pragma unroll
                            for (int u = 0; u < 8; u++) cmac(cr, ci, va[u], vb[u]);
                        }
                        for (; s + 3 * G < KT; s += 4 * G) {
                            double2 va[4], vb[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) { va[u] = pah[ws.sumA[s + u * G]]; vb[u] = pbh[ws.sumB[s + u * G]]; }
#pragma unroll
                            for (int u = 0; u < 4; u++) cmac(cr, ci, va[u], vb[u]);
                        }
                        for (; s < KT; s += G) cmac(cr, ci, pah[ws.sumA[s]], pbh[ws.sumB[s]]);
                    }
                }
                for (uint32_t off = P; off < 32; off <<= 1) {            // no-op when G = 1
                    cr += __shfl_xor_sync(0xffffffffu, cr, off);
                    ci += __shfl_xor_sync(0xffffffffu, ci, off);
                }
                if (c < NC && sg == 0) st.C[c] = make_double2(cr, ci);
            }
        }
        if (csize > 1) {
            // release / acquire at cluster scope covers the level's global-memory writes (a gpu-scope __threadfence() in front of it
            // was 8 % of the stall samples, profiles/r02_ncu_summary.txt, and is not needed: all readers are in this cluster)
            asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
        } else {
            __syncthreads();
        }
        if (timeline && threadIdx.x == 0 && crank == 0) timeline[lvl + 1] = clock64();
    }
}

static constexpr auto k_micro = k_micro_t<QTB_MICRO_THREADS>;                 // single plans / eager groups
static constexpr auto k_micro_batch = k_micro_t<QTB_MICRO_THREADS_BATCH>;     // term batches (launch_micro_plans)

}  // namespace qtb
