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
// Structure (one persistent CTA per SM, 8 warps):
//   * tile gather  : 16-byte cp.async.cg (LDGSTS) per element, lanes ordered along the operand's
//                    memory-contiguous bits (host-computed bit permutation) -> >= 64 B runs always
//   * pipeline     : STAGES-deep ring over the FLAT (tile, k-chunk) sequence, so the next tile's
//                    operands stream in while the current tile's DMMAs and C stores are in flight
//   * math         : mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no f64 kind; this
//                    is the B200 FP64 tensor pipe, 37.1 TFLOP/s measured), complex product as
//                    4 real DMMAs (Xr*Yr, -Xi*Yi, Xr*Yi, Xi*Yr); fragments come out of shared
//                    memory as one conflict-free LDS.128 (re, im) per lane
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t scatter_bits(uint32_t v, const uint8_t *sh, int first, int count) {
    uint32_t o = 0;
    for (int j = 0; j < count; j++) o += ((v >> j) & 1u) << sh[first + j];
    return o;
}

template <int WX, int WY, int FX, int FY, int TK, int STAGES>
struct GettCfg {
    static constexpr int NT = WX * WY * 32;
    static constexpr int TM = WX * FX * 8, TN = WY * FY * 8;
    static constexpr int LDX = TM + 2, LDY = TN + 2;          // +2 (x16 B): conflict-free LDS.128 fragments
    static constexpr int XS = TK * LDX, YS = TK * LDY;        // elements per stage
    static constexpr int STAGE_ELEMS = XS + YS;
    static constexpr int TAB = 2 * TM + 2 * TN + 2 * TK + 64; // uint32 tables
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * 16 + TAB * 4;
    static constexpr int XSLOTS = (TM * TK + NT - 1) / NT, YSLOTS = (TN * TK + NT - 1) / NT;
};

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

template <int WX, int WY, int FX, int FY, int TK, int STAGES>
__global__ void __launch_bounds__(WX *WY * 32, 1) k_gett(const GettParams p) {
    using Cfg = GettCfg<WX, WY, FX, FY, TK, STAGES>;
    constexpr int NT = Cfg::NT, TM = Cfg::TM, TN = Cfg::TN, LDX = Cfg::LDX, LDY = Cfg::LDY;
    constexpr int TMB = ilog2(TM), TNB = ilog2(TN), TKB = ilog2(TK);
    constexpr int XSLOTS = Cfg::XSLOTS, YSLOTS = Cfg::YSLOTS;

    extern __shared__ __align__(16) uint8_t smemRaw[];
    double2 *stages = reinterpret_cast<double2 *>(smemRaw);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smemRaw + (size_t)STAGES * Cfg::STAGE_ELEMS * 16);
    uint32_t *tXx = tab, *tCx = tXx + TM, *tYy = tCx + TM, *tCy = tYy + TN, *tXk = tCy + TN, *tYk = tXk + TK;
    uint32_t *dXo = tYk + TK, *dXs = dXo + 16, *dYo = dXs + 16, *dYs = dYo + 16;   // per-slot-round deltas

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wx0 = (warp % WX) * (FX * 8), wy0 = (warp / WX) * (FY * 8);
    const uint32_t nyValid = p.nyValid;

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
    // zero the operand ring once: padded y columns (N < TN) are never written again
    for (int i = tid; i < STAGES * Cfg::STAGE_ELEMS; i += NT) stages[i] = make_double2(0.0, 0.0);
    __syncthreads();

    // slot id e = tid + r*NT  ->  tile coordinate via the load-order permutation; the contribution
    // of the r bits is CTA-uniform, so only the tid part lives in registers.
    auto coordOf = [&](uint32_t e, const uint8_t *perm, int nbits) {
        uint32_t w = 0;
        for (int j = 0; j < nbits; j++) w |= ((e >> j) & 1u) << perm[j];
        return w;                                       // low bits: x (or y) local, bits >= TMB/TNB: k local
    };
    constexpr int XEB = TMB + TKB;
    const int yeb = p.nyBits + TKB;                     // valid Y-tile slot bits
    const uint32_t nYElems = 1u << yeb;
    uint32_t xOff0, xSm0, yOff0 = 0, ySm0 = 0;
    {
        const uint32_t w = coordOf(tid & ((1u << XEB) - 1), p.permX, XEB);
        const uint32_t xl = w & (TM - 1), kl = w >> TMB;
        xOff0 = tXx[xl] + tXk[kl];
        xSm0 = (kl * LDX + xl) * 16;
        const uint32_t wy = coordOf(tid & (nYElems - 1), p.permY, yeb);
        const uint32_t yl = wy & (TN - 1), kly = wy >> TNB;
        yOff0 = tYy[yl] + tYk[kly];
        ySm0 = (Cfg::XS + kly * LDY + yl) * 16;
    }
    if (tid < XSLOTS) {
        const uint32_t w = coordOf((uint32_t)tid * NT & ((1u << XEB) - 1), p.permX, XEB);
        const uint32_t xl = w & (TM - 1), kl = w >> TMB;
        dXo[tid] = tXx[xl] + tXk[kl];
        dXs[tid] = (kl * LDX + xl) * 16;
    }
    if (tid < YSLOTS) {
        const uint32_t e = (uint32_t)tid * NT;
        const uint32_t wy = e < nYElems ? coordOf(e, p.permY, yeb) : 0u;
        const uint32_t yl = wy & (TN - 1), kly = wy >> TNB;
        dYo[tid] = tYy[yl] + tYk[kly];
        dYs[tid] = (kly * LDY + yl) * 16;
    }
    __syncthreads();

    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(stages);
    const uint32_t nTiles = p.nTilesX * p.nTilesY, nChunks = p.nChunks;
    const uint32_t myTiles = blockIdx.x < nTiles ? (nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total = myTiles * nChunks;

    auto issue = [&](uint32_t q) {
        // flat pipeline index q -> (tile, chunk)
        const uint32_t ti = q / nChunks, ch = q - ti * nChunks;
        const uint32_t tile = blockIdx.x + ti * gridDim.x;
        const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
        const uint32_t stage = q % STAGES;
        const uint32_t chX = scatter_bits(ch, p.shXk, TKB, p.kbits - TKB);
        const uint32_t chY = scatter_bits(ch, p.shYk, TKB, p.kbits - TKB);
        const double2 *gx = p.X + scatter_bits(tx, p.shXx, TMB, p.xbits - TMB) + chX + xOff0;
        const double2 *gy = p.Y + scatter_bits(ty, p.shYy, p.nyBits, p.ybits - p.nyBits) + chY + yOff0;
        const uint32_t sb = smemBase + stage * (Cfg::STAGE_ELEMS * 16);
#pragma unroll
        for (int r = 0; r < XSLOTS; r++) {
            if (XSLOTS * NT == TM * TK || tid + r * NT < TM * TK) cp_async16(sb + xSm0 + dXs[r], gx + dXo[r]);
        }
#pragma unroll
        for (int r = 0; r < YSLOTS; r++) {
            if ((uint32_t)(tid + r * NT) < nYElems) cp_async16(sb + ySm0 + dYs[r], gy + dYo[r]);
        }
    };

    // NOTE on the slot decomposition: slot e = tid + r*NT with NT a power of two, so the bits of e
    // split into the tid bits and the r bits; offsets are sums over disjoint bit contributions, hence
    // off(e) = off(tid) + off(r*NT).  When the tile has fewer than NT elements the tid is masked.

    double accR[FX][FY][2], accI[FX][FY][2];

#pragma unroll 1
    for (uint32_t q = 0; q < (uint32_t)(STAGES - 1); q++) {
        if (q < total) issue(q);
        cp_async_commit();
    }

#pragma unroll 1
    for (uint32_t q = 0; q < total; q++) {
        const uint32_t ti = q / nChunks, ch = q - ti * nChunks;
        if (ch == 0) {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++) { accR[i][j][0] = accR[i][j][1] = accI[i][j][0] = accI[i][j][1] = 0.0; }
        }
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (q + STAGES - 1 < total) issue(q + STAGES - 1);
        cp_async_commit();

        const double2 *xs = stages + (size_t)(q % STAGES) * Cfg::STAGE_ELEMS;
        const double2 *ys = xs + Cfg::XS;
#pragma unroll
        for (int kk = 0; kk < TK / 4; kk++) {
            double2 xf[FX], yf[FY];
#pragma unroll
            for (int i = 0; i < FX; i++) xf[i] = xs[(kk * 4 + t) * LDX + wx0 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < FY; j++) yf[j] = ys[(kk * 4 + t) * LDY + wy0 + j * 8 + g];
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

        if (ch == nChunks - 1) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            double2 *cb = p.C + scatter_bits(tx, p.shCx, TMB, p.xbits - TMB) + scatter_bits(ty, p.shCy, p.nyBits, p.ybits - p.nyBits);
#pragma unroll
            for (int i = 0; i < FX; i++) {
                const uint32_t ox = tCx[wx0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < FY; j++) {
                    const int y0 = wy0 + j * 8 + 2 * t;
                    if ((uint32_t)y0 < nyValid) cb[ox + tCy[y0]] = make_double2(accR[i][j][0], accI[i][j][0]);
                    if ((uint32_t)(y0 + 1) < nyValid) cb[ox + tCy[y0 + 1]] = make_double2(accR[i][j][1], accI[i][j][1]);
                }
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace qtb
