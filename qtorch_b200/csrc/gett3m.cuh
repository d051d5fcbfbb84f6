// qtorch_b200/csrc/gett3m.cuh -- the compute-bound class of Network::ContractIndices (/root/reference/src/Network.h:892-935)
// on the FP64 tensor pipe with the 3M (Karatsuba) complex product and SHARED operand sums (sm_100a).
//
// OPT-IN (QTB_GETT_C1=5): correct, but measured 5 % slower than k_gett's 3M variant on the rank-14 steps of BASELINE config 2
// (3.68 against 3.51 ms; why: profiles/r02_g3_shared_sums.txt), so the default stays k_gett.
//
// Same contract as k_gett (gett.cuh): D[x, y] = sum_k X[x, k] * Y[k, y] on the operands' original layouts, 64 x 64 x 16
// tiles, 16 math warps of 16 x 16 (a 4 x 4 grid) + one producer warpgroup, cp.async gather into an mbarrier ring.  What differs:
//
//   * 3M needs S = re + im of every operand element.  k_gett lets every math warp add them for its own fragments, so each
//     sum is formed four times per tile (once per warp that shares the row / column block) by FP64 adds that compete with
//     the DMMAs for the FP64 pipe: 64 DADD per 192 DMMA, measured 10 % of the step (profiles/r01_dadd_tax.txt).  Here each
//     sum is formed ONCE: the X row block wx is shared by the four warps (wx, 0..3) and a ring slot has four k-steps, so warp
//     (wx, wy) adds the fragments of k-step wy of block wx -- and, symmetrically, k-step wx of Y column block wy -- for the
//     slot AFTER the one it is about to multiply, stores them (STS.128) into a small ring of sum tiles and arrives on that
//     tile's `summed` mbarrier.  16 DADD per slot and thread become 4; the DMMA loop reads sums with LDS.128 next to the
//     (re, im) fragments and has no FP64 add left.  A sum tile is stored fragment-native ([block][k-step][lane][i]), so
//     writer and readers use the same lane and no index arithmetic.  (Letting the PRODUCER warpgroup form the sums was
//     measured first and dropped: a DADD issued by a fifth warp queues ~325 cycles behind the math warps' DMMAs, the sums
//     arrive late: 3.79-3.93 ms against 3.52, profiles/r02_g3_shared_sums.txt.)
//   * ordering: a warp multiplies slot q after `summed[q % 3]` completes, i.e. after all 16 warps have started slot q - 1;
//     the tile written at slot q (for q + 1) was last read at slot q - 2, which every warp has left by then.
//   * operand tiles are unpadded and XOR-swizzled (32 KB per stage instead of 40 KB), which is what makes room for three sum
//     tiles next to a 5-deep operand ring inside 227 KB.  Both layouts keep LDS.128 fragment reads (quarter-warps)
//     conflict-free:
//         k contiguous in HBM:  slot(f, k) = 16 f + (k ^ ((f & 3) << 2))
//         f contiguous in HBM:  slot(f, k) = 64 k + (f ^ ((k & 3) << 1))                 (f = x or y inside the tile)
//     Every slot function is linear over GF(2), so a gather slot's shared-memory offset is off(thread) XOR off(round), as
//     the global offset is off(thread) + off(round).
#pragma once
#include "gett.cuh"

namespace qtb {

template <int STAGES, int SD, int FUSE>
struct Gett3Cfg {
    static constexpr int NW = 16, NPT = 128, NT = NW * 32 + NPT;
    static constexpr int TM = 64, TN = 64, TK = 16;
    static constexpr int TILE = TM * TK;                       // elements of one operand tile
    static constexpr int STAGE_ELEMS = 2 * TILE;               // X tile, Y tile: 32 KB
    static constexpr int ROUNDS = TILE / NPT;                  // gather rounds per operand tile
    static constexpr int MATH_REGS = 104, PROD_REGS = 64;
    static constexpr int LA = 1;                               // sums are formed this many slots ahead; SD = 2 LA + 1 sum tiles
    // fused inner product: the matching dotD tile (64 x 64) rides through the ring in DP pieces of DROWS rows, [y][x] layout
    static constexpr int DROWS = FUSE ? STAGE_ELEMS / TM : 0, DP = FUSE ? TN / (DROWS ? DROWS : 1) : 0;
    static constexpr int DROUNDS = FUSE ? DROWS * TM / NPT : 0;
    static constexpr int TAB = 2 * TM + 2 * TN + 2 * TK + 4 * ROUNDS + (FUSE ? TM + TN + 2 * DROUNDS : 0);
    static constexpr int NBARS = 2 * STAGES + SD;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * 16 + (size_t)SD * STAGE_ELEMS * 8 + TAB * 4 + 16 + NW * 16 * FUSE + NBARS * 8 + 16 +
                                   (6 + 2 * FUSE) * sizeof(HiTab);
    static_assert(SMEM <= 232448, "k_gett3s: shared memory over the 227 KB per-CTA limit");
};

__device__ __forceinline__ uint32_t g3_tile_slot(uint32_t f, uint32_t k, bool kFast) {
    return kFast ? (f * 16u + (k ^ ((f & 3u) << 2))) : (k * 64u + (f ^ ((k & 3u) << 1)));
}

template <int STAGES, int SD, int FUSE>
__global__ void __launch_bounds__(640, 1) k_gett3s(const GettParams p) {
    using Cfg = Gett3Cfg<STAGES, SD, FUSE>;
    constexpr int NW = Cfg::NW, NT = Cfg::NT, TM = Cfg::TM, TN = Cfg::TN, TK = Cfg::TK, ROUNDS = Cfg::ROUNDS;
    constexpr int TMB = 6, TNB = 6, TKB = 4, FX = 2, FY = 2;
    constexpr uint32_t STAGE_BYTES = Cfg::STAGE_ELEMS * 16, SUM_BYTES = Cfg::STAGE_ELEMS * 8;

    extern __shared__ __align__(16) uint8_t smemRaw[];
    double2 *stages = reinterpret_cast<double2 *>(smemRaw);
    double *sums = reinterpret_cast<double *>(smemRaw + (size_t)STAGES * STAGE_BYTES);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smemRaw + (size_t)STAGES * STAGE_BYTES + (size_t)SD * SUM_BYTES);
    uint32_t *tXx = tab, *tCx = tXx + TM, *tYy = tCx + TM, *tCy = tYy + TN, *tXk = tCy + TN, *tYk = tXk + TK;
    uint32_t *dXo = tYk + TK, *dXs = dXo + ROUNDS, *dYo = dXs + ROUNDS, *dYs = dYo + ROUNDS;
    uint32_t *tDx = dYs + ROUNDS, *tDy = tDx + (FUSE ? TM : 0);
    uint32_t *dDo = tDy + (FUSE ? TN : 0), *dDs = dDo + Cfg::DROUNDS;
    double2 *dotRed = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(dDs + Cfg::DROUNDS) + 15) & ~(uintptr_t)15);
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(FUSE ? (uint32_t *)(dotRed + NW) : dDs) + 7) & ~(uintptr_t)7);
    HiTab *hi = reinterpret_cast<HiTab *>((reinterpret_cast<uintptr_t>(bars + Cfg::NBARS) + 15) & ~(uintptr_t)15);
    const uint32_t barBase = (uint32_t)__cvta_generic_to_shared(bars);     // full[s] = +8 s, empty[s] = +8 (STAGES + s), summed[j] = +8 (2 STAGES + j)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool xK = p.xsK == 1, yK = p.ysK == 1;                 // operand tile layout: k (true) or x / y (false) contiguous

    for (int i = tid; i < TM; i += NT) {
        tXx[i] = scatter_bits(i, p.shXx, 0, TMB);
        tCx[i] = scatter_bits(i, p.shCx, 0, TMB);
        tYy[i] = scatter_bits(i, p.shYy, 0, TNB);
        tCy[i] = scatter_bits(i, p.shCy, 0, TNB);
        if (FUSE) { tDx[i] = scatter_bits(i, p.shDx, 0, TMB); tDy[i] = scatter_bits(i, p.shDy, 0, TNB); }
    }
    for (int i = tid; i < TK; i += NT) {
        tXk[i] = scatter_bits(i, p.shXk, 0, TKB);
        tYk[i] = scatter_bits(i, p.shYk, 0, TKB);
    }
    const int xHiBits = p.xbits - TMB, yHiBits = p.ybits - TNB, kHiBits = p.kbits - TKB;
    const int xParts = (xHiBits + 6) / 7, yParts = (yHiBits + 6) / 7, kParts = (kHiBits + 6) / 7;
    hitab_build(hi[0], p.shXx, TMB, xHiBits, tid, NT);
    hitab_build(hi[1], p.shYy, TNB, yHiBits, tid, NT);
    hitab_build(hi[2], p.shXk, TKB, kHiBits, tid, NT);
    hitab_build(hi[3], p.shYk, TKB, kHiBits, tid, NT);
    hitab_build(hi[4], p.shCx, TMB, xHiBits, tid, NT);
    hitab_build(hi[5], p.shCy, TNB, yHiBits, tid, NT);
    if (FUSE) {
        hitab_build(hi[6], p.shDx, TMB, xHiBits, tid, NT);
        hitab_build(hi[7], p.shDy, TNB, yHiBits, tid, NT);
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(barBase + 8 * s, Cfg::NPT); mbar_init(barBase + 8 * (STAGES + s), NW); }
        for (int j = 0; j < SD; j++) mbar_init(barBase + 8 * (2 * STAGES + j), NW);
    }
    __syncthreads();

    auto coordOf = [&](uint32_t e, const uint8_t *perm, int nbits) {
        uint32_t w = 0;
        for (int j = 0; j < nbits; j++) w |= ((e >> j) & 1u) << perm[j];
        return w;                                       // low 6 bits: x (or y) inside the tile, bits >= 6: k inside the chunk
    };
    constexpr int EB = TMB + TKB;
    constexpr int DEB = FUSE ? TMB + ilog2(Cfg::DROWS ? Cfg::DROWS : 1) : 0;
    // dotD piece, [y][x] with 64-element rows; x ^ (((y >> 1) & 3) << 1) keeps the math warps' reads (rows 2t + h) apart
    auto dSlot = [](uint32_t xl, uint32_t yl) { return yl * 64u + (xl ^ (((yl >> 1) & 3u) << 1)); };
    for (int r = tid; r < ROUNDS; r += NT) {
        const uint32_t w = coordOf((uint32_t)r << 7, p.permX, EB);
        dXo[r] = tXx[w & (TM - 1)] + tXk[w >> TMB];
        dXs[r] = g3_tile_slot(w & (TM - 1), w >> TMB, xK) * 16;
        const uint32_t wy = coordOf((uint32_t)r << 7, p.permY, EB);
        dYo[r] = tYy[wy & (TN - 1)] + tYk[wy >> TNB];
        dYs[r] = g3_tile_slot(wy & (TN - 1), wy >> TNB, yK) * 16;
    }
    if (FUSE) {
        for (int r = tid; r < Cfg::DROUNDS; r += NT) {
            const uint32_t w = coordOf((uint32_t)r << 7, p.permD, DEB);
            dDo[r] = tDx[w & (TM - 1)] + tDy[w >> TMB];
            dDs[r] = dSlot(w & (TM - 1), w >> TMB) * 16;
        }
    }
    __syncthreads();

    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(stages);
    const uint32_t nTiles = p.nTilesX * p.nTilesY, nChunks = p.nChunks;
    const uint32_t myTiles = blockIdx.x < nTiles ? (nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t perTile = nChunks + (FUSE ? Cfg::DP : 0);          // ring slots per tile: k-chunks, then dotD pieces
    const uint32_t total = myTiles * perTile;

    if (warp >= NW) {
        // ================= producer warpgroup =================
        setmaxnreg_dec<Cfg::PROD_REGS>();
        const uint32_t ptid = tid - NW * 32;
        uint32_t xOff0, xSm0, yOff0, ySm0, dOff0 = 0, dSm0 = 0;
        {
            const uint32_t w = coordOf(ptid, p.permX, EB);
            xOff0 = tXx[w & (TM - 1)] + tXk[w >> TMB];
            xSm0 = g3_tile_slot(w & (TM - 1), w >> TMB, xK) * 16;
            const uint32_t wy = coordOf(ptid, p.permY, EB);
            yOff0 = tYy[wy & (TN - 1)] + tYk[wy >> TNB];
            ySm0 = g3_tile_slot(wy & (TN - 1), wy >> TNB, yK) * 16;
            if (FUSE) {
                const uint32_t wd = coordOf(ptid, p.permD, DEB);
                dOff0 = tDx[wd & (TM - 1)] + tDy[wd >> TMB];
                dSm0 = dSlot(wd & (TM - 1), wd >> TMB) * 16;
            }
        }
        uint32_t tiI = 0, chI = 0;                                     // (tile, slot-in-tile) of the next slot to issue
        for (uint32_t q = 0; q < total; q++) {
            const uint32_t stage = q % STAGES;
            if (q >= (uint32_t)STAGES) mbar_wait(barBase + 8 * (STAGES + stage), (q / STAGES - 1) & 1);
            const uint32_t tile = blockIdx.x + tiI * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            const uint32_t sb = smemBase + stage * STAGE_BYTES;
            if (FUSE && chI >= nChunks) {
                const uint32_t piece = chI - nChunks;
                const double2 *gd = p.dotD + hitab_lookup(hi[6], tx, xParts) + hitab_lookup(hi[7], ty, yParts) + tDy[piece * Cfg::DROWS] + dOff0;
#pragma unroll 8
                for (int r = 0; r < Cfg::DROUNDS; r++) cp_async16(sb + (dSm0 ^ dDs[r]), gd + dDo[r]);
            } else {
                const double2 *gx = p.X + hitab_lookup(hi[0], tx, xParts) + hitab_lookup(hi[2], chI, kParts) + xOff0;
                const double2 *gy = p.Y + hitab_lookup(hi[1], ty, yParts) + hitab_lookup(hi[3], chI, kParts) + yOff0;
#pragma unroll
                for (int r = 0; r < ROUNDS; r++) cp_async16(sb + (xSm0 ^ dXs[r]), gx + dXo[r]);
#pragma unroll
                for (int r = 0; r < ROUNDS; r++) cp_async16(sb + Cfg::TILE * 16 + (ySm0 ^ dYs[r]), gy + dYo[r]);
            }
            cp_async_mbar_arrive(barBase + 8 * stage);
            if (++chI == perTile) { chI = 0; ++tiI; }
        }
        cp_async_wait<0>();
        return;
    }

    // ================= math warps =================
    setmaxnreg_inc<Cfg::MATH_REGS>();
    const int g = lane >> 2, t = lane & 3;
    const int wx0 = (warp & 3) * 16, wy0 = (warp >> 2) * 16;
    // fragment addressing (byte offsets inside a stage; see the header):  e(i, kk) = ((E0 + i * Ia) ^ (kk * Kx)) + kk * Ka
    const int bx = warp & 3, by = warp >> 2;
    const uint32_t xE0 = xK ? (uint32_t)((wx0 + g) * 16 + t + 4 * (g & 3)) : (uint32_t)(t * 64 + wx0 + (g ^ (2 * t)));
    const uint32_t yE0 = yK ? (uint32_t)((wy0 + g) * 16 + t + 4 * (g & 3)) : (uint32_t)(t * 64 + wy0 + (g ^ (2 * t)));
    const uint32_t xKx = xK ? 64u : 0u, xKa = xK ? 0u : 4096u, xIa = xK ? 128u : 8u;       // bytes
    const uint32_t yKx = yK ? 64u : 0u, yKa = yK ? 0u : 4096u, yIa = yK ? 128u : 8u;
    uint32_t xE[FX], yE[FY];                            // the kk = 0 fragments
#pragma unroll
    for (int i = 0; i < FX; i++) xE[i] = (xE0 + i * xIa) * 16;
#pragma unroll
    for (int j = 0; j < FY; j++) yE[j] = (Cfg::TILE + yE0 + j * yIa) * 16;
    // sum tiles, fragment-native: X part [block 0..3][k-step 0..3][lane][i], then the Y part; 16 bytes per lane and (block, k-step)
    const uint32_t sumBaseM = (uint32_t)__cvta_generic_to_shared(sums);
    const uint32_t xSr = ((bx * 4) * 32 + lane) * 16, ySr = 8192u + ((by * 4) * 32 + lane) * 16;        // readers: + kk * 512
    const uint32_t xSw = xSr + by * 512, ySw = ySr + bx * 512;                                          // this warp's share: k-step by of X, bx of Y
    const uint32_t xDuty0 = (xE[0] ^ (by * xKx)) + by * xKa, xDuty1 = (xE[1] ^ (by * xKx)) + by * xKa;
    const uint32_t yDuty0 = (yE[0] ^ (bx * yKx)) + bx * yKa, yDuty1 = (yE[1] ^ (bx * yKx)) + bx * yKa;

    auto ldsE = [](uint32_t addr) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v; };
    auto stsE = [](uint32_t addr, double a, double b) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory"); };

    // this warp's share of the operand sums of ring slot s (X block bx at k-step by, Y block by at k-step bx), done at the start
    // of slot s - 1.  (Placing the adds inside the DMMA burst of slot s - 1 instead -- half a share per k-step, arrival after
    // k-step 1 -- measured slower: 3.80-4.04 ms against 3.68-3.70.)  dutyChunk: slot s is a k-chunk slot (a dotD piece has no sums).
    uint32_t chD = 0;
    auto addv = [](double a, double b) { double r; asm volatile("add.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r; };
    // half h = 0: the X share, h = 1: the Y share (one half per k-step keeps the transient registers at two fragments)
    auto dutyLoad = [&](uint32_t s, int h, double2 &d0, double2 &d1) {
        const uint32_t eb = smemBase + (s % STAGES) * STAGE_BYTES;
        d0 = ldsE(eb + (h ? yDuty0 : xDuty0)); d1 = ldsE(eb + (h ? yDuty1 : xDuty1));
    };
    auto dutyStore = [&](uint32_t s, int h, const double2 &d0, const double2 &d1) {
        stsE(sumBaseM + (s % SD) * SUM_BYTES + (h ? ySw : xSw), addv(d0.x, d0.y), addv(d1.x, d1.y));
    };
    auto dutyArrive = [&](uint32_t s) {
        __syncwarp();
        if (lane == 0) mbar_arrive(barBase + 8 * (2 * STAGES + s % SD));
        if (++chD == perTile) chD = 0;
    };

    double accR[FX][FY][2], accI[FX][FY][2], acc3[FX][FY][2];          // T1 = Xr Yr, T2 = Xi Yi, T3 = Xs Ys
    double dotR = 0.0, dotI = 0.0;
    uint32_t ti = 0, ch = 0;
    if (total > 0) {                                                           // slot 0 is always a k-chunk slot
        mbar_wait(barBase, 0);
        double2 d0, d1;
        dutyLoad(0, 0, d0, d1); dutyStore(0, 0, d0, d1);
        dutyLoad(0, 1, d0, d1); dutyStore(0, 1, d0, d1);
        dutyArrive(0);
    }
#pragma unroll 1
    for (uint32_t q = 0; q < total; q++) {
        if (ch == 0) {
#pragma unroll
            for (int i = 0; i < FX; i++)
#pragma unroll
                for (int j = 0; j < FY; j++)
                    accR[i][j][0] = accR[i][j][1] = accI[i][j][0] = accI[i][j][1] = acc3[i][j][0] = acc3[i][j][1] = 0.0;
        }
        const uint32_t stage = q % STAGES, sj = q % SD;
        // all 16 shares of this slot's sums are in (=> a k-chunk slot has landed), and every warp has left slot q - 2, whose sum
        // tile the share for slot q + 1 overwrites
        mbar_wait(barBase + 8 * (2 * STAGES + sj), (q / SD) & 1);
        const bool hasDuty = q + 1 < total, dutyChunk = !FUSE || chD < nChunks;
        if (hasDuty) {
            if (dutyChunk) {
                mbar_wait(barBase + 8 * ((q + 1) % STAGES), ((q + 1) / STAGES) & 1);
                double2 d0, d1;
                dutyLoad(q + 1, 0, d0, d1); dutyStore(q + 1, 0, d0, d1);
                dutyLoad(q + 1, 1, d0, d1); dutyStore(q + 1, 1, d0, d1);
            }
            dutyArrive(q + 1);
        }
        if (!FUSE || ch < nChunks) {
            const uint32_t eb = smemBase + stage * STAGE_BYTES, sb = sumBaseM + sj * SUM_BYTES;
            double2 xf[2][FX], yf[2][FY], xs[2], ys[2];                        // xs[.] = (sum i = 0, sum i = 1)
#pragma unroll
            for (int i = 0; i < FX; i++) xf[0][i] = ldsE(eb + xE[i]);
#pragma unroll
            for (int j = 0; j < FY; j++) yf[0][j] = ldsE(eb + yE[j]);
            xs[0] = ldsE(sb + xSr); ys[0] = ldsE(sb + ySr);
#pragma unroll
            for (int kk = 0; kk < TK / 4; kk++) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < TK / 4) {
#pragma unroll
                    for (int i = 0; i < FX; i++) xf[nxt][i] = ldsE(eb + (xE[i] ^ ((kk + 1) * xKx)) + (kk + 1) * xKa);
#pragma unroll
                    for (int j = 0; j < FY; j++) yf[nxt][j] = ldsE(eb + (yE[j] ^ ((kk + 1) * yKx)) + (kk + 1) * yKa);
                    xs[nxt] = ldsE(sb + xSr + (kk + 1) * 512); ys[nxt] = ldsE(sb + ySr + (kk + 1) * 512);
                }
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accR[i][j][0], accR[i][j][1], xf[cur][i].x, yf[cur][j].x);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(accI[i][j][0], accI[i][j][1], xf[cur][i].y, yf[cur][j].y);
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++) dmma884(acc3[i][j][0], acc3[i][j][1], i ? xs[cur].y : xs[cur].x, j ? ys[cur].y : ys[cur].x);
            }
        } else {
            // dotD piece in this stage: the warps whose rows it holds fold it into the inner product
            const uint32_t piece = ch - nChunks;
            mbar_wait(barBase + 8 * stage, (q / STAGES) & 1);
            if ((uint32_t)wy0 / (uint32_t)(Cfg::DROWS ? Cfg::DROWS : 1) == piece) {
                const double2 *dt = stages + (size_t)stage * Cfg::STAGE_ELEMS;
                const uint32_t yBase = (uint32_t)wy0 - piece * Cfg::DROWS;
#pragma unroll
                for (int i = 0; i < FX; i++)
#pragma unroll
                    for (int j = 0; j < FY; j++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const uint32_t yl = yBase + j * 8 + 2 * t + h, xl = wx0 + i * 8 + g;
                            const double2 dv = dt[yl * 64u + (xl ^ (((yl >> 1) & 3u) << 1))];
                            const double re = accR[i][j][h] - accI[i][j][h], im = acc3[i][j][h] - accR[i][j][h] - accI[i][j][h];
                            dotR = fma(re, dv.x, dotR); dotR = fma(-im, dv.y, dotR);
                            dotI = fma(re, dv.y, dotI); dotI = fma(im, dv.x, dotI);
                        }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(barBase + 8 * (STAGES + stage));

        if (!FUSE && ch == nChunks - 1) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t ty = tile / p.nTilesX, tx = tile - ty * p.nTilesX;
            double2 *cb = p.C + hitab_lookup(hi[4], tx, xParts) + hitab_lookup(hi[5], ty, yParts);
#pragma unroll
            for (int i = 0; i < FX; i++) {
                const uint32_t ox = tCx[wx0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < FY; j++) {
                    const int y0 = wy0 + j * 8 + 2 * t;
                    cb[ox + tCy[y0]] = make_double2(accR[i][j][0] - accI[i][j][0], acc3[i][j][0] - accR[i][j][0] - accI[i][j][0]);
                    cb[ox + tCy[y0 + 1]] = make_double2(accR[i][j][1] - accI[i][j][1], acc3[i][j][1] - accR[i][j][1] - accI[i][j][1]);
                }
            }
        }
        if (++ch == perTile) { ch = 0; ++ti; }
    }
    if (FUSE) {
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
