// qtorch_b200/csrc/gett.cuh -- fused index-permute + contract on the FP64 tensor pipe (sm_100a).
//
// Replaces the big steps of Network::ContractIndices (/root/reference/src/Network.h:892-935):
//      D[x, y] = sum_k X[x, k] * Y[k, y]        (complex<double>)
// where X, Y are the two Node tensors in their ORIGINAL leg order and x / y / k are arbitrary
// subsets of legs ("GETT": the transposition is folded into the global->shared tile gather, no
// permuted copy of an operand is ever written to HBM).  Which Node plays X is the host's choice
// (the operand with the larger free dimension), so D is C or C^T; every operand and the output are
// addressed through per-bit shift tables, so any leg placement is legal.
//
// Structure (one persistent CTA per SM: 16 or 8 math warps + a producer warpgroup, mbarrier full/empty ring):
//   * tile gather  : the producer warpgroup issues 16-byte cp.async.cg (LDGSTS) per element, lanes ordered along the
//                    operand's memory-contiguous bits (host-computed bit permutation) -> >= 64 B runs always; the
//                    dimension that is contiguous in HBM is also contiguous in shared memory, so both halves of a
//                    32-byte sector are served by one request
//   * pipeline     : STAGES-deep ring over the FLAT (tile, k-chunk) sequence, so the next tile's operands stream in
//                    while the current tile's DMMAs and C stores are in flight; cp.async.mbarrier.arrive.noinc signals
//                    "full", one arrive per math warp signals "empty"; setmaxnreg shifts registers to the math warps
//   * math         : mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no f64 kind; this is the B200 FP64
//                    tensor pipe, 37.1 TFLOP/s measured); complex product as 4 real DMMAs (4M) or 3 (3M, Karatsuba);
//                    fragments come out of shared memory as one conflict-free LDS.128 (re, im) per lane
//   * epilogue     : accumulators -> C directly, 128-byte runs per row group
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qtb {

struct GettParams {
    const double2 *X;
    const double2 *Y;
    double2 *C;
    uint32_t nTilesX, nTilesY, nChunks;
    uint32_t nyValid;                 // valid y extent inside a tile (== TN unless N < TN)
    uint8_t xbits, ybits, kbits;      // total index bits (2 per leg)
    uint8_t nyBits;                   // log2(nyValid)
    uint8_t shXx[32], shCx[32];       // x bit j -> bit position in X's / C's element index
    uint8_t shYy[32], shCy[32];       // y bit j -> bit position in Y's / C's element index
    uint8_t shXk[32], shYk[32];       // k bit j -> bit position in X's / Y's element index
    uint8_t permX[16];                // load order: bit j of the X-tile slot id -> tile coord bit
    uint8_t permY[16];                //   (coord id < TMB: x/y bit, else TMB + k bit)
    // shared-memory element strides of the two tiles.  The tile dimension that is contiguous in HBM is also made
    // contiguous in shared memory: the two 16-byte halves of a 32-byte sector then land side by side and one
    // LDGSTS sector request serves both lanes (otherwise every sector is requested twice).  Both layouts give
    // conflict-free LDS.128 fragment reads: [k][x] with row stride T+2, or [x][k] with row stride TK+4 (== 4 mod 8).
    uint16_t xsX, xsK, ysY, ysK;
    // fused inner product (FUSE = 1 instantiations): instead of writing D, accumulate sum D[x,y] * dotD[perm(x,y)] --
    // the step that follows contracts the result with `dotD` over ALL legs, so the 4^rC-element intermediate is never
    // written to or re-read from HBM.  shDx/shDy: x / y bit j -> bit position inside dotD's element index.
    const double2 *dotD;
    double2 *dotPartial;              // one partial sum per CTA
    uint8_t shDx[32], shDy[32];
    uint8_t permD[16];                // load order of a dotD tile piece: slot bit j -> coordinate bit (x bits, then the piece's y bits)
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dneg(const double v) {
    return __longlong_as_double(__double_as_longlong(v) ^ (long long)0x8000000000000000ull);   // integer pipe
}
__device__ __forceinline__ void cp_async16(uint32_t smemAddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr), "l"(g) : "memory");
}
// one 32-byte sector per store (sm_100 STG.256): two adjacent complex results of a lane
__device__ __forceinline__ void st_global_256(double2 *dst, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t scatter_bits(uint32_t v, const uint8_t *sh, int first, int count) {
    uint32_t o = 0;
    for (int j = 0; j < count; j++) o += ((v >> j) & 1u) << sh[first + j];
    return o;
}

// High (per-tile / per-chunk) index bits -> element offset through 7-bit lookup tables in shared memory: a
// bit-by-bit scatter costs ~4 instructions per bit, which made the streaming tile shapes issue-bound.
#define QTB_HT_PARTS 4
#define QTB_HT_SIZE 128
struct HiTab { uint32_t t[QTB_HT_PARTS][QTB_HT_SIZE]; };
__device__ __forceinline__ void hitab_build(HiTab &tab, const uint8_t *sh, int first, int count, int tid, int nthreads) {
    for (int idx = tid; idx < QTB_HT_PARTS * QTB_HT_SIZE; idx += nthreads) {
        const int part = idx / QTB_HT_SIZE, v = idx % QTB_HT_SIZE;
        const int lo = 7 * part, n = count - lo < 0 ? 0 : (count - lo > 7 ? 7 : count - lo);
        tab.t[part][v] = scatter_bits((uint32_t)v, sh, first + lo, n);
    }
}
__device__ __forceinline__ uint32_t hitab_lookup(const HiTab &tab, uint32_t v, int parts) {
    uint32_t o = tab.t[0][v & 127u];
    if (parts > 1) o += tab.t[1][(v >> 7) & 127u];
    if (parts > 2) o += tab.t[2][(v >> 14) & 127u];
    if (parts > 3) o += tab.t[3][(v >> 21) & 127u];
    return o;
}

template <int WX, int WY, int FX, int FY, int TK, int STAGES, int MODE3M = 0, int FUSE = 0>
struct GettCfg {
    static constexpr int NW = WX * WY;                        // math warps; one more warpgroup (4 warps) produces
    static constexpr int NPT = 128;                           // producer threads
    static constexpr int NT = NW * 32 + NPT;
    // register split (setmaxnreg): the per-scheduler total must stay within what the CTA was launched with
    // (8 math warps: 3 warps/scheduler x 168 = 504 = 2 x 224 + 56;  16 math warps: 5 x 96 = 480 = 4 x 104 + 64)
    static constexpr int MATH_REGS = NW == 8 ? 224 : 104, PROD_REGS = NW == 8 ? 56 : 64;
    static constexpr int TM = WX * FX * 8, TN = WY * FY * 8;
    static constexpr int LDX = TM + 2, LDY = TN + 2;          // [k][x] layout: +2 (x16 B) keeps LDS.128 fragments conflict-free
    static constexpr int LDK = (TK % 8 == 0) ? TK + 4 : TK;   // [x][k] layout: row stride == 4 (mod 8)
    static constexpr int XS = (TK * LDX > TM * LDK) ? TK * LDX : TM * LDK;        // elements per stage (either layout fits)
    static constexpr int YS = (TK * LDY > TN * LDK) ? TK * LDY : TN * LDK;
    static constexpr int STAGE_ELEMS = XS + YS;
    static constexpr int XROUNDS = (TM * TK + NPT - 1) / NPT, YROUNDS = (TN * TK + NPT - 1) / NPT;
    // fused inner product: the matching dotD tile (TM x TN) rides through the same ring after a tile's last k-chunk,
    // cut into DP pieces of TN/DP rows ([y][x] layout, row stride TM+2) so that each piece fits one stage
    static constexpr int dpFor(int dp) { return (TN / dp) * (TM + 2) <= STAGE_ELEMS ? dp : dpFor(dp * 2); }
    static constexpr int DP = FUSE ? dpFor(1) : 0;
    static constexpr int DROWS = FUSE ? TN / (DP ? DP : 1) : 0;                 // y rows per piece
    static constexpr int DROUNDS = FUSE ? (DROWS * TM + NPT - 1) / NPT : 0;
    static constexpr int TAB = 2 * TM + 2 * TN + 2 * TK + 2 * XROUNDS + 2 * YROUNDS + (FUSE ? TM + TN + 64 + 2 * DROUNDS : 0);   // uint32 tables
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * 16 + TAB * 4 + 2 * STAGES * 8 + 16 + 32 * FUSE + (6 + 2 * FUSE) * sizeof(HiTab);
};

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// ---- mbarrier helpers (CTA scope) ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE;\n"
        " bra WAIT_LOOP;\n"
        "DONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
// arrive on `bar` once every cp.async issued so far by this thread has landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Warp-specialised: warps 0..NW-1 do LDS + DMMA + the C stores, the last warpgroup only gathers tiles (cp.async)
// and signals "full" mbarriers; math warps hand stages back through "empty" mbarriers.  The math pipe never waits
// for address arithmetic or load issue, and the gather of the next tiles proceeds under the epilogue stores.
// MODE3M = 1: the complex product uses three real DMMAs per tile pair instead of four (Karatsuba / "3M":
//   T1 = Xr*Yr, T2 = Xi*Yi, T3 = (Xr+Xi)*(Yr+Yi);  Re = T1 - T2, Im = T3 - T1 - T2), 25 % less tensor-pipe work for
// the same algorithmic 8*U flops, at the price of a third accumulator set.
template <int WX, int WY, int FX, int FY, int TK, int STAGES, int MODE3M = 0, int FUSE = 0>
__global__ void __launch_bounds__(WX *WY * 32 + 128, 1) k_gett(const GettParams p) {
    using Cfg = GettCfg<WX, WY, FX, FY, TK, STAGES, MODE3M, FUSE>;
    constexpr int NW = Cfg::NW, NT = Cfg::NT, TM = Cfg::TM, TN = Cfg::TN;
    const uint32_t xsX = p.xsX, xsK = p.xsK, ysY = p.ysY, ysK = p.ysK;
    constexpr int TMB = ilog2(TM), TNB = ilog2(TN), TKB = ilog2(TK);
    constexpr int XROUNDS = Cfg::XROUNDS, YROUNDS = Cfg::YROUNDS;

    extern __shared__ __align__(16) uint8_t smemRaw[];
    double2 *stages = reinterpret_cast<double2 *>(smemRaw);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smemRaw + (size_t)STAGES * Cfg::STAGE_ELEMS * 16);
    uint32_t *tXx = tab, *tCx = tXx + TM, *tYy = tCx + TM, *tCy = tYy + TN, *tXk = tCy + TN, *tYk = tXk + TK;
    uint32_t *dXo = tYk + TK, *dXs = dXo + XROUNDS, *dYo = dXs + XROUNDS, *dYs = dYo + YROUNDS;   // per-round deltas
    uint32_t *tDx = dYs + YROUNDS, *tDy = tDx + (FUSE ? TM : 0);                                   // fused dot: tile-local -> dotD offset
    uint32_t *dDo = tDy + (FUSE ? TN : 0), *dDs = dDo + Cfg::DROUNDS;                                // dotD piece: per-round deltas
    double2 *dotRed = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(dDs + Cfg::DROUNDS) + 15) & ~(uintptr_t)15);   // NW partials
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(FUSE ? (uint32_t *)(dotRed + NW) : dYs + YROUNDS) + 7) & ~(uintptr_t)7);
    HiTab *hi = reinterpret_cast<HiTab *>(bars + 2 * STAGES);     // [0] X by tile-x, [1] Y by tile-y, [2] X by chunk, [3] Y by chunk, [4] C by tile-x, [5] C by tile-y
    const uint32_t barBase = (uint32_t)__cvta_generic_to_shared(bars);        // full[s] = barBase + 8 s, empty[s] = full + 8 STAGES

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nyValid = p.nyValid;

    // Programmatic dependent launch (opt-in, QTB_PDL=1: the host then sets cudaLaunchAttributeProgrammaticStreamSerialization; both
    // instructions are no-ops otherwise): the
    // NEXT tile-kernel launch of the stream may start its CTAs on every SM this grid has left and build its tables while this grid's
    // last tiles run; it waits (griddepcontrol.wait below) for this grid to complete before it touches global memory.
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- one-time tables: tile-local coordinate -> element offset
    for (int i = tid; i < TM; i += NT) {
        tXx[i] = scatter_bits(i, p.shXx, 0, TMB);
        tCx[i] = scatter_bits(i, p.shCx, 0, TMB);
    }
    for (int i = tid; i < TN; i += NT) {
        const int nb = p.nyBits;
        tYy[i] = (uint32_t)i < nyValid ? scatter_bits(i, p.shYy, 0, nb) : 0u;
        tCy[i] = (uint32_t)i < nyValid ? scatter_bits(i, p.shCy, 0, nb) : 0u;
    }
    for (int i = tid; i < TK; i += NT) {
        tXk[i] = scatter_bits(i, p.shXk, 0, TKB);
        tYk[i] = scatter_bits(i, p.shYk, 0, TKB);
    }
    const int xHiBits = p.xbits - TMB, yHiBits = p.ybits - p.nyBits, kHiBits = p.kbits - TKB;
    const int xParts = (xHiBits + 6) / 7, yParts = (yHiBits + 6) / 7, kParts = (kHiBits + 6) / 7;
    hitab_build(hi[0], p.shXx, TMB, xHiBits, tid, NT);
    hitab_build(hi[1], p.shYy, p.nyBits, yHiBits, tid, NT);
    hitab_build(hi[2], p.shXk, TKB, kHiBits, tid, NT);
    hitab_build(hi[3], p.shYk, TKB, kHiBits, tid, NT);
    hitab_build(hi[4], p.shCx, TMB, xHiBits, tid, NT);
    hitab_build(hi[5], p.shCy, p.nyBits, yHiBits, tid, NT);
    if (FUSE) {
        hitab_build(hi[6], p.shDx, TMB, xHiBits, tid, NT);
        hitab_build(hi[7], p.shDy, p.nyBits, yHiBits, tid, NT);
        for (int i = tid; i < TM; i += NT) tDx[i] = scatter_bits(i, p.shDx, 0, TMB);
        for (int i = tid; i < TN; i += NT) tDy[i] = (uint32_t)i < nyValid ? scatter_bits(i, p.shDy, 0, p.nyBits) : 0u;
    }
    // zero the operand ring once when the tile has padded y columns (N < TN): they are never written again
    if (nyValid < (uint32_t)TN) for (int i = tid; i < STAGES * Cfg::STAGE_ELEMS; i += NT) stages[i] = make_double2(0.0, 0.0);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(barBase + 8 * s, Cfg::NPT); mbar_init(barBase + 8 * (STAGES + s), NW); }
    }
    __syncthreads();

    // slot id e = ptid + 128*r  ->  tile coordinate via the load-order permutation (bits follow the operand's
    // memory significance); offsets are sums over disjoint bit contributions, so off(e) = off(lane) + off(32 r)
    auto coordOf = [&](uint32_t e, const uint8_t *perm, int nbits) {
        uint32_t w = 0;
        for (int j = 0; j < nbits; j++) w |= ((e >> j) & 1u) << perm[j];
        return w;                                       // low bits: x (or y) local, bits >= TMB/TNB: k local
    };
    constexpr int XEB = TMB + TKB;
    const int yeb = p.nyBits + TKB;                     // valid Y-tile slot bits
    const uint32_t nYElems = 1u << yeb;
    for (int r = tid; r < XROUNDS; r += NT) {
        const uint32_t w = coordOf((uint32_t)r << 7, p.permX, XEB);
        const uint32_t xl = w & (TM - 1), kl = w >> TMB;
        dXo[r] = tXx[xl] + tXk[kl];
        dXs[r] = (kl * xsK + xl * xsX) * 16;
    }
    for (int r = tid; r < YROUNDS; r += NT) {
        const uint32_t e = (uint32_t)r << 7;
        const uint32_t wy = e < nYElems ? coordOf(e, p.permY, yeb) : 0u;
        const uint32_t yl = wy & (TN - 1), kly = wy >> TNB;
        dYo[r] = tYy[yl] + tYk[kly];
        dYs[r] = (kly * ysK + yl * ysY) * 16;
    }
    constexpr int DEB = FUSE ? TMB + ilog2(Cfg::DROWS ? Cfg::DROWS : 1) : 0;        // slot bits of one dotD piece
    if (FUSE) {
        for (int r = tid; r < Cfg::DROUNDS; r += NT) {
            const uint32_t w = coordOf((uint32_t)r << 7, p.permD, DEB);
            const uint32_t xl = w & (TM - 1), yl = w >> TMB;
            dDo[r] = tDx[xl] + tDy[yl];
            dDs[r] = (yl * (TM + 2) + xl) * 16;
        }
    }
    __syncthreads();

    // everything above touched parameters and shared memory only; operands and C belong to the grids before this one
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(stages);
    const uint32_t nTiles = p.nTilesX * p.nTilesY, nChunks = p.nChunks;
    const uint32_t myTiles = blockIdx.x < nTiles ? (nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t perTile = nChunks + (FUSE ? Cfg::DP : 0);          // ring slots per tile: k-chunks, then dotD pieces
    const uint32_t total = myTiles * perTile;

    if (warp >= NW) {
        // ================= producer warpgroup =================
        setmaxnreg_dec<Cfg::PROD_REGS>();
        const uint32_t ptid = tid - NW * 32;
        uint32_t xOff0, xSm0, yOff0, ySm0;
        {
            const uint32_t w = coordOf(ptid & ((1u << XEB) - 1), p.permX, XEB);
            const uint32_t xl = w & (TM - 1), kl = w >> TMB;
            xOff0 = tXx[xl] + tXk[kl];
            xSm0 = (kl * xsK + xl * xsX) * 16;
            const uint32_t wy = coordOf(ptid & (nYElems - 1), p.permY, yeb);
            const uint32_t yl = wy & (TN - 1), kly = wy >> TNB;
            yOff0 = tYy[yl] + tYk[kly];
            ySm0 = (Cfg::XS + kly * ysK + yl * ysY) * 16;
        }
        const int yRounds = (int)((nYElems + Cfg::NPT - 1) / Cfg::NPT);
        const bool yLane = ptid < nYElems;
        const bool xLane = ptid < (uint32_t)(TM * TK);
        uint32_t dOff0 = 0, dSm0 = 0;
        if (FUSE) {
            const uint32_t w = coordOf(ptid & ((1u << DEB) - 1), p.permD, DEB);
            const uint32_t xl = w & (TM - 1), yl = w >> TMB;
            dOff0 = tDx[xl] + tDy[yl];
            dSm0 = (yl * (TM + 2) + xl) * 16;
        }
        uint32_t ti = 0, ch = 0;
        for (uint32_t q = 0; q < total; q++) {
            const uint32_t stage = q % STAGES, round = q / STAGES;
            if (round > 0) mbar_wait(barBase + 8 * (STAGES + stage), (round - 1) & 1);
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            if (FUSE && ch >= nChunks) {
                // dotD piece (ch - nChunks): rows [piece*DROWS, (piece+1)*DROWS) of this tile, gathered in dotD's memory order
                const uint32_t piece = ch - nChunks;
                const double2 *gd = p.dotD + hitab_lookup(hi[6], tx, xParts) + hitab_lookup(hi[7], ty, yParts) + tDy[piece * Cfg::DROWS] + dOff0;
                const uint32_t sbd = smemBase + stage * (Cfg::STAGE_ELEMS * 16);
#pragma unroll 8
                for (int r = 0; r < Cfg::DROUNDS; r++) cp_async16(sbd + dSm0 + dDs[r], gd + dDo[r]);
                cp_async_mbar_arrive(barBase + 8 * stage);
                if (++ch == perTile) { ch = 0; ++ti; }
                continue;
            }
            const double2 *gx = p.X + hitab_lookup(hi[0], tx, xParts) + hitab_lookup(hi[2], ch, kParts) + xOff0;
            const double2 *gy = p.Y + hitab_lookup(hi[1], ty, yParts) + hitab_lookup(hi[3], ch, kParts) + yOff0;
            const uint32_t sb = smemBase + stage * (Cfg::STAGE_ELEMS * 16);
            if (xLane) {
#pragma unroll 8
                for (int r = 0; r < XROUNDS; r++) cp_async16(sb + xSm0 + dXs[r], gx + dXo[r]);
            }
            if (yLane) {
#pragma unroll 4
                for (int r = 0; r < yRounds; r++) cp_async16(sb + ySm0 + dYs[r], gy + dYo[r]);
            }
            cp_async_mbar_arrive(barBase + 8 * stage);
            if (++ch == perTile) { ch = 0; ++ti; }
        }
        cp_async_wait<0>();
        return;
    }

    // ================= math warps =================
    setmaxnreg_inc<Cfg::MATH_REGS>();
    const int g = lane >> 2, t = lane & 3;
    const int wx0 = (warp % WX) * (FX * 8), wy0 = (warp / WX) * (FY * 8);
    const uint32_t xFrag0 = (wx0 + g) * xsX + t * xsK, yFrag0 = (wy0 + g) * ysY + t * ysK;     // this lane's fragment origin
    // 4M: accR = Re, accI = Im.   3M: accR = T1, accI = T2, acc3 = T3.
    double accR[FX][FY][2], accI[FX][FY][2], acc3[MODE3M ? FX : 1][MODE3M ? FY : 1][2];
    double dotR = 0.0, dotI = 0.0;                      // FUSE: running sum D[x,y] * dotD[..] of this thread
    uint32_t ti = 0, ch = 0;
#pragma unroll 1
    for (uint32_t q = 0; q < total; q++) {
        if (ch == 0) {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) {
                    accR[i][j][0] = accR[i][j][1] = accI[i][j][0] = accI[i][j][1] = 0.0;
                    if (MODE3M) acc3[i][j][0] = acc3[i][j][1] = 0.0;
                }
        }
        const uint32_t stage = q % STAGES;
        mbar_wait(barBase + 8 * stage, (q / STAGES) & 1);
        if (!FUSE || ch < nChunks) {
        const double2 *xs_ = stages + (size_t)stage * Cfg::STAGE_ELEMS;
        const double2 *ys_ = xs_ + Cfg::XS;
        if (MODE3M) {
            // software-pipelined: the fragments (and their re+im sums, which need the FP64 pipe the DMMAs saturate)
            // of k-step kk+1 are fetched while the DMMAs of k-step kk issue, so no DADD sits on the critical path
            double2 xf[2][FX], yf[2][FY];
            double xs[2][FX], ys[2][FY];
#pragma unroll
            for (int i = 0; i < FX; i++) { xf[0][i] = xs_[xFrag0 + i * 8 * xsX]; xs[0][i] = xf[0][i].x + xf[0][i].y; }
#pragma unroll
            for (int j = 0; j < FY; j++) { yf[0][j] = ys_[yFrag0 + j * 8 * ysY]; ys[0][j] = yf[0][j].x + yf[0][j].y; }
#pragma unroll
            for (int kk = 0; kk < TK / 4; kk++) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < TK / 4) {
#pragma unroll
                    for (int i = 0; i < FX; i++) xf[nxt][i] = xs_[xFrag0 + (kk + 1) * 4 * xsK + i * 8 * xsX];
#pragma unroll
                    for (int j = 0; j < FY; j++) yf[nxt][j] = ys_[yFrag0 + (kk + 1) * 4 * ysK + j * 8 * ysY];
                }
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accR[i][j][0], accR[i][j][1], xf[cur][i].x, yf[cur][j].x);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accI[i][j][0], accI[i][j][1], xf[cur][i].y, yf[cur][j].y);
                if (kk + 1 < TK / 4) {
#pragma unroll
                    for (int i = 0; i < FX; i++) xs[nxt][i] = xf[nxt][i].x + xf[nxt][i].y;
#pragma unroll
                    for (int j = 0; j < FY; j++) ys[nxt][j] = yf[nxt][j].x + yf[nxt][j].y;
                }
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(acc3[i][j][0], acc3[i][j][1], xs[cur][i], ys[cur][j]);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < TK / 4; kk++) {
                double2 xf[FX], yf[FY];
#pragma unroll
                for (int i = 0; i < FX; i++) xf[i] = xs_[xFrag0 + kk * 4 * xsK + i * 8 * xsX];
#pragma unroll
                for (int j = 0; j < FY; j++) yf[j] = ys_[yFrag0 + kk * 4 * ysK + j * 8 * ysY];
                // four passes over the FX x FY accumulator tiles, one per real product: consecutive DMMAs never
                // touch the same accumulator (a dependent pair is FX*FY issues apart)
                double nxi[FX];
#pragma unroll
                for (int i = 0; i < FX; i++) nxi[i] = dneg(xf[i].y);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accR[i][j][0], accR[i][j][1], xf[i].x, yf[j].x);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accI[i][j][0], accI[i][j][1], xf[i].x, yf[j].y);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accR[i][j][0], accR[i][j][1], nxi[i], yf[j].y);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accI[i][j][0], accI[i][j][1], xf[i].y, yf[j].x);
            }
        }
        } else {
            // dotD piece in this stage ([y][x], row stride TM+2): the warps whose rows it holds fold it into the inner product
            const uint32_t piece = ch - nChunks;
            if ((uint32_t)wy0 / (uint32_t)(Cfg::DROWS ? Cfg::DROWS : 1) == piece) {
                const double2 *dt = stages + (size_t)stage * Cfg::STAGE_ELEMS;
                const uint32_t yBase = (uint32_t)wy0 - piece * Cfg::DROWS;
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const double2 dv = dt[(yBase + j * 8 + 2 * t + h) * (TM + 2) + wx0 + i * 8 + g];
                            double re, im;
                            if (MODE3M) { re = accR[i][j][h] - accI[i][j][h]; im = acc3[i][j][h] - accR[i][j][h] - accI[i][j][h]; }
                            else { re = accR[i][j][h]; im = accI[i][j][h]; }
                            dotR = fma(re, dv.x, dotR); dotR = fma(-im, dv.y, dotR);
                            dotI = fma(re, dv.y, dotI); dotI = fma(im, dv.x, dotI);
                        }
            }
        }
        // stage consumed: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(barBase + 8 * (STAGES + stage));

        if (!FUSE && ch == nChunks - 1) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            double2 *cb = p.C + hitab_lookup(hi[4], tx, xParts) + hitab_lookup(hi[5], ty, yParts);
            // D = C^T (the y legs come first in C): a lane's two results y0, y0 + 1 are neighbours in memory and fill one
            // 32-byte sector -- written with one 256-bit store instead of two half-sector stores (streaming tile shapes
            // only: there the store stream is the bound, 1.23 -> 0.94 ms for (3,13,k=1); the compute-bound 64x64 / 128x64
            // configurations keep the plain epilogue, whose schedule the extra path disturbed: 3.52 -> 3.68 ms)
            constexpr bool PAIRABLE = (WY == 1);
            const bool pairStore = PAIRABLE && nyValid >= 2 && p.shCy[0] == 0;
#pragma unroll
            for (int i = 0; i < FX; i++) {
                const uint32_t ox = tCx[wx0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < FY; j++) {
                    const int y0 = wy0 + j * 8 + 2 * t;
                    double re0, im0, re1, im1;
                    if (MODE3M) {
                        re0 = accR[i][j][0] - accI[i][j][0]; im0 = acc3[i][j][0] - accR[i][j][0] - accI[i][j][0];
                        re1 = accR[i][j][1] - accI[i][j][1]; im1 = acc3[i][j][1] - accR[i][j][1] - accI[i][j][1];
                    } else {
                        re0 = accR[i][j][0]; im0 = accI[i][j][0]; re1 = accR[i][j][1]; im1 = accI[i][j][1];
                    }
                    if (PAIRABLE && pairStore) {
                        if ((uint32_t)y0 < nyValid) st_global_256(cb + ox + tCy[y0], re0, im0, re1, im1);
                    } else {
                        if ((uint32_t)y0 < nyValid) cb[ox + tCy[y0]] = make_double2(re0, im0);
                        if ((uint32_t)(y0 + 1) < nyValid) cb[ox + tCy[y0 + 1]] = make_double2(re1, im1);
                    }
                }
            }
        }
        if (++ch == perTile) { ch = 0; ++ti; }
    }
    if (FUSE) {
        // CTA partial of the fused inner product: shuffle tree per warp, then a named barrier over the math warps only
        // (the producer warpgroup may already have exited)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { dotR += __shfl_xor_sync(0xffffffffu, dotR, o); dotI += __shfl_xor_sync(0xffffffffu, dotI, o); }
        if (lane == 0) dotRed[warp] = make_double2(dotR, dotI);
        asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
        if (tid == 0) {
            double sr = 0.0, si = 0.0;
            for (int w = 0; w < NW; w++) { sr += dotRed[w].x; si += dotRed[w].y; }
            p.dotPartial[blockIdx.x] = make_double2(sr, si);
        }
    }
}

}  // namespace qtb
