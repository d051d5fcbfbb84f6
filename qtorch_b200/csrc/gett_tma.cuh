// qtorch_b200/csrc/gett_tma.cuh -- the DMMA tile kernel with TMA-fed operand tiles (sm_100a).
//
// Same contraction as gett.cuh -- D[x, y] = sum_k X[x, k] * Y[k, y] on the operands' ORIGINAL layouts
// (Network::ContractIndices, /root/reference/src/Network.h:892-935) -- but an operand tile is not gathered element by
// element: ONE elected thread issues one cp.async.bulk.tensor (TMA, SASS UTMALDG) per operand and ring stage, completion
// is signalled on the stage's mbarrier by transaction bytes (mbarrier.expect_tx), and the 128 producer threads of
// gett.cuh with their per-element address arithmetic disappear.
//
// How a base-4 digit gather becomes a TMA box.  A rank-r tensor is 4^r elements of 16 bytes, leg 0 fastest; a tile takes
// a few legs whole (3 free legs + 2 shared legs for a 64 x 16 tile) at the digits of all other legs (tile index, k-chunk
// index), and always holds leg 0 (the host's tile choice guarantees it).  The tensor map has dim0 = leg 0 x (re, im)
// [8 x f64 = 64 contiguous bytes], dim1 = an "offset" dimension of 4^(r-1) units with a 64-byte stride and box extent 1 --
// its coordinate is the element offset of the tile's corner / 4 -- and one dimension per RUN of memory-adjacent tile legs
// above leg 0 (size 4^len, stride 16 * 4^pos bytes, box = full size).  The dimensions alias each other in memory, which
// the driver accepts (tools/probe_tma.cu, profiles/r02_probe_tma.txt).  Up to three run dimensions fit (at most four legs
// each); the host checks that and may pick other free legs for the tile to get there.
//
// Shared-memory layout.  TMA writes the box densely in memory order of the tile's legs; in a swizzled mode every inner row
// is padded to the swizzle span (probe: 16-byte rows under the 128-byte swizzle land 128 bytes apart), so the inner row is
// made exactly one 64-byte span and the 64-byte swizzle is used: 16-byte slot s = idx ^ ((idx >> 3) & 3).  The DMMA
// fragment of a quarter-warp is 8 lanes = (row bit g0, k bits t0, t1); it is bank-conflict free iff those three coordinate
// bits sit at dense-index positions whose swizzle vectors are linearly independent (positions 0-2 -> e_p, 3-4 -> e_(p-3),
// >= 5 -> nothing).  Which tile bit plays g0 / t0 / t1 is free -- the logical x / y / k numbering is only a labelling that
// the C-address tables absorb -- so the host searches an assignment that satisfies BOTH operands (t0, t1 are the same
// shared-leg bits in X and Y) and hands the kernel, per logical bit, its dense-index position.  The swizzle is linear over
// GF(2): a fragment address is laneBase ^ c_i[i] ^ c_kk[kk] with the two constants in the parameter bank.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "gett.cuh"

namespace qtb {

struct alignas(64) GettTmaParams {
    CUtensorMap tmX, tmY;             // see above; box = one TM x TK (TN x TK) tile
    GettParams g;                     // leg tables (C addressing, tile / chunk offsets), sizes
    uint8_t ixX[8], ikX[8];           // dense-index bit position of logical x bit b / logical k bit b inside an X tile
    uint8_t iyY[8], ikY[8];           // ... of logical y bit b / k bit b inside a Y tile
    uint16_t cXi[4], cXkk[8];         // swizzled byte offsets contributed by the fragment index i (x += 8 i) / k-step kk
    uint16_t cYj[4], cYkk[8];
};

template <int WX, int WY, int FX, int FY, int TK, int STAGES, int MODE3M>
struct GettTmaCfg {
    static constexpr int NW = WX * WY;
    static constexpr int NT = NW * 32 + 32;                    // math warps + ONE producer warp (a single lane issues the TMA)
    static constexpr int TM = WX * FX * 8, TN = WY * FY * 8;
    static constexpr int XBYTES = TM * TK * 16, YBYTES = TN * TK * 16;
    static constexpr int STAGE_BYTES = XBYTES + YBYTES;
    static constexpr int TAB = 2 * TM + 2 * TN;                // tCx, tCy (+ spare): uint32
    static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + TAB * 4 + 2 * STAGES * 8 + 16 + 6 * sizeof(HiTab);
};

__device__ __forceinline__ uint32_t swz64_bytes(uint32_t idx) { return (idx ^ ((idx >> 3) & 3u)) << 4; }

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *tm, int32_t off, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(0), "r"(off), "r"(0), "r"(0), "r"(0), "r"(bar) : "memory");
}

template <int WX, int WY, int FX, int FY, int TK, int STAGES, int MODE3M>
__global__ void __launch_bounds__(WX *WY * 32 + 32, 1) k_gett_tma(const __grid_constant__ GettTmaParams P) {
    using Cfg = GettTmaCfg<WX, WY, FX, FY, TK, STAGES, MODE3M>;
    constexpr int NW = Cfg::NW, NT = Cfg::NT, TM = Cfg::TM, TN = Cfg::TN;
    constexpr int TMB = ilog2(TM), TNB = ilog2(TN), TKB = ilog2(TK);
    const GettParams &p = P.g;

    extern __shared__ uint8_t smemRawUnaligned[];
    // the swizzled TMA destination wants 1024-byte alignment
    uint8_t *smemRaw = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smemRawUnaligned) + 1023) & ~(uintptr_t)1023);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smemRaw + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint32_t *tCx = tab, *tCy = tCx + TM;
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(tCy + TN + TM + TN) + 7) & ~(uintptr_t)7);
    HiTab *hi = reinterpret_cast<HiTab *>(bars + 2 * STAGES);     // [0] X by tile-x, [1] Y by tile-y, [2] X by chunk, [3] Y by chunk, [4] C by tile-x, [5] C by tile-y
    const uint32_t barBase = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(smemRaw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < TM; i += NT) tCx[i] = scatter_bits(i, p.shCx, 0, TMB);
    for (int i = tid; i < TN; i += NT) tCy[i] = scatter_bits(i, p.shCy, 0, TNB);
    const int xHiBits = p.xbits - TMB, yHiBits = p.ybits - TNB, kHiBits = p.kbits - TKB;
    const int xParts = (xHiBits + 6) / 7, yParts = (yHiBits + 6) / 7, kParts = (kHiBits + 6) / 7;
    hitab_build(hi[0], p.shXx, TMB, xHiBits, tid, NT);
    hitab_build(hi[1], p.shYy, TNB, yHiBits, tid, NT);
    hitab_build(hi[2], p.shXk, TKB, kHiBits, tid, NT);
    hitab_build(hi[3], p.shYk, TKB, kHiBits, tid, NT);
    hitab_build(hi[4], p.shCx, TMB, xHiBits, tid, NT);
    hitab_build(hi[5], p.shCy, TNB, yHiBits, tid, NT);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(barBase + 8 * s, 1); mbar_init(barBase + 8 * (STAGES + s), NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nTiles = p.nTilesX * p.nTilesY, nChunks = p.nChunks;
    const uint32_t myTiles = blockIdx.x < nTiles ? (nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total = myTiles * nChunks;

    if (warp >= NW) {
        // ================= producer warp: lane 0 feeds X tiles, lane 1 feeds Y tiles, one TMA box each per ring stage =====
        // (lane 0 also posts the stage's expected byte count; a box that lands before that only drives the barrier's
        // transaction count negative for a moment -- the phase cannot complete before the arrive.expect_tx)
        if (lane < 2) {
            const CUtensorMap *tm = lane == 0 ? &P.tmX : &P.tmY;
            asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
            const HiTab &hTile = lane == 0 ? hi[0] : hi[1], &hChunk = lane == 0 ? hi[2] : hi[3];
            const int tParts = lane == 0 ? xParts : yParts;
            const uint32_t sub = lane == 0 ? 0u : (uint32_t)Cfg::XBYTES;
            uint32_t ti = 0, ch = 0;
            for (uint32_t q = 0; q < total; q++) {
                const uint32_t stage = q % STAGES, round = q / STAGES;
                if (round > 0) mbar_wait(barBase + 8 * (STAGES + stage), (round - 1) & 1);
                const uint32_t tile = blockIdx.x + ti * gridDim.x;
                const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
                const uint32_t off = hitab_lookup(hTile, lane == 0 ? tx : ty, tParts) + hitab_lookup(hChunk, ch, kParts);
                const uint32_t full = barBase + 8 * stage;
                if (lane == 0) mbar_expect_tx(full, Cfg::STAGE_BYTES);
                tma_load_5d(smemBase + stage * Cfg::STAGE_BYTES + sub, tm, (int32_t)(off >> 2), full);    // offset dimension: units of leg 0
                if (++ch == nChunks) { ch = 0; ++ti; }
            }
        }
        return;
    }

    // ================= math warps =================
    const int g = lane >> 2, t = lane & 3;
    const int wx0 = (warp % WX) * (FX * 8), wy0 = (warp / WX) * (FY * 8);
    // this lane's fragment origin inside a stage (swizzled byte offsets; the swizzle is linear, see the header)
    uint32_t laneX, laneY;
    {
        uint32_t ix = 0, iy = 0;
        const uint32_t xl = (uint32_t)(wx0 + g), yl = (uint32_t)(wy0 + g);
#pragma unroll
        for (int b = 0; b < TMB; b++) ix |= ((xl >> b) & 1u) << P.ixX[b];
#pragma unroll
        for (int b = 0; b < TNB; b++) iy |= ((yl >> b) & 1u) << P.iyY[b];
#pragma unroll
        for (int b = 0; b < 2; b++) { ix |= (((uint32_t)t >> b) & 1u) << P.ikX[b]; iy |= (((uint32_t)t >> b) & 1u) << P.ikY[b]; }
        laneX = swz64_bytes(ix);
        laneY = swz64_bytes(iy) + Cfg::XBYTES;
    }
    double accR[FX][FY][2], accI[FX][FY][2], acc3[MODE3M ? FX : 1][MODE3M ? FY : 1][2];
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
        const uint32_t st_ = smemBase + stage * Cfg::STAGE_BYTES;
        auto ldX = [&](int i, int kk) { return lds_f64x2(st_ + (laneX ^ (uint32_t)P.cXi[i] ^ (uint32_t)P.cXkk[kk])); };
        auto ldY = [&](int j, int kk) { return lds_f64x2(st_ + (laneY ^ (uint32_t)P.cYj[j] ^ (uint32_t)P.cYkk[kk])); };
        if (MODE3M) {
            double2 xf[2][FX], yf[2][FY];
            double xs[2][FX], ys[2][FY];
#pragma unroll
            for (int i = 0; i < FX; i++) { xf[0][i] = ldX(i, 0); xs[0][i] = xf[0][i].x + xf[0][i].y; }
#pragma unroll
            for (int j = 0; j < FY; j++) { yf[0][j] = ldY(j, 0); ys[0][j] = yf[0][j].x + yf[0][j].y; }
#pragma unroll
            for (int kk = 0; kk < TK / 4; kk++) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < TK / 4) {
#pragma unroll
                    for (int i = 0; i < FX; i++) xf[nxt][i] = ldX(i, kk + 1);
#pragma unroll
                    for (int j = 0; j < FY; j++) yf[nxt][j] = ldY(j, kk + 1);
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
                for (int i = 0; i < FX; i++) xf[i] = ldX(i, kk);
#pragma unroll
                for (int j = 0; j < FY; j++) yf[j] = ldY(j, kk);
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
        // stage consumed: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(barBase + 8 * (STAGES + stage));

        if (ch == nChunks - 1) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            double2 *cb = p.C + hitab_lookup(hi[4], tx, xParts) + hitab_lookup(hi[5], ty, yParts);
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
                    cb[ox + tCy[y0]] = make_double2(re0, im0);
                    cb[ox + tCy[y0 + 1]] = make_double2(re1, im1);
                }
            }
        }
        if (++ch == nChunks) { ch = 0; ++ti; }
    }
}

}  // namespace qtb
