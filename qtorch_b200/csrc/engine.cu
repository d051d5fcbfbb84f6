// qtorch_b200/csrc/engine.cu -- libqtorch_b200.so: the C ABI declared in include/qtorch_b200.h.
//
// Device-resident tensor storage (per-rank pooled: every tensor is exactly 16*4^rank bytes, so one
// free list per rank is a perfect stream-ordered allocator), the step dispatcher that maps each
// Network::ContractIndices call (/root/reference/src/Network.h:876) to a kernel family, the grouped
// micro-step executor, compiled plans (CUDA-graph backed) and the NCCL scalar reduction.
//
// No CPU fallback exists: without a CUDA device every compute entry returns QTB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/qtorch_b200.h"
#include "step.h"
#include "kernels.cuh"
#include "gett.cuh"
#include "gett_tma.cuh"
#include "gett3m.cuh"
#include "reduce.cuh"
#include "apply.cuh"

using namespace qtb;

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_lastError;
static int fail(int status, const std::string &msg) { g_lastError = msg; return status; }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? QTB_ERR_OOM : QTB_ERR_CUDA,              \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)
#define ST(call) do { int s_ = (call); if (s_ != QTB_OK) return s_; } while (0)

// ------------------------------------------------------------------------------------------------
// kernel families
// A micro-step runs on ONE SM inside the grouped launch, so its size limit trades launch count against the
// load bandwidth of a single SM (2 x 16 B per complex MAC): 4^6 MACs by default for latency-bound single plans,
// raised to 4^8..4^10 by callers that run many plans side by side (qtb_ctx_set_micro_limit; the QAOA term dispatcher).
static const int MICRO_DEFAULT_LOG4 = 6;
static const int MICRO_MAX_RANK = 7;

struct GettChoice { int cfg; bool swap; };      // cfg: 0 C1/TK16 1 C1/TK4 2 C2/TK16 3 C2/TK4 4 C3/TK16 5 C3/TK4

typedef void (*GettKernel)(const GettParams);
struct GettInst { GettKernel fn; int TM, TN, TK, NT; size_t smem; int occ; int drows; };   // drows: dotD rows per ring stage (fused variants)
// C1: 128x64 tile, compute-bound big x big;  C2: 256x16;  C3: 256x8 (N <= 4 padded) -- streaming
static GettInst g_gett[15] = {
    {k_gett<4, 2, 4, 4, 16, 3>, 128, 64, 16, GettCfg<4, 2, 4, 4, 16, 3>::NT, GettCfg<4, 2, 4, 4, 16, 3>::SMEM, 1},
    {k_gett<4, 2, 4, 4, 4, 12>, 128, 64, 4, GettCfg<4, 2, 4, 4, 4, 12>::NT, GettCfg<4, 2, 4, 4, 4, 12>::SMEM, 1},
    {k_gett<8, 1, 4, 2, 16, 2>, 256, 16, 16, GettCfg<8, 1, 4, 2, 16, 2>::NT, GettCfg<8, 1, 4, 2, 16, 2>::SMEM, 1},
    {k_gett<8, 1, 4, 2, 4, 10>, 256, 16, 4, GettCfg<8, 1, 4, 2, 4, 10>::NT, GettCfg<8, 1, 4, 2, 4, 10>::SMEM, 1},
    {k_gett<8, 1, 4, 1, 16, 2>, 256, 8, 16, GettCfg<8, 1, 4, 1, 16, 2>::NT, GettCfg<8, 1, 4, 1, 16, 2>::SMEM, 1},
    {k_gett<8, 1, 4, 1, 4, 10>, 256, 8, 4, GettCfg<8, 1, 4, 1, 4, 10>::NT, GettCfg<8, 1, 4, 1, 4, 10>::SMEM, 1},
    // C1 with 16 math warps of 32x16
    {k_gett<4, 4, 4, 2, 16, 3>, 128, 64, 16, GettCfg<4, 4, 4, 2, 16, 3>::NT, GettCfg<4, 4, 4, 2, 16, 3>::SMEM, 1},
    // C1 "3M": 64x64 tile, 16 math warps of 16x16, three real DMMAs per complex tile product
    {k_gett<4, 4, 2, 2, 16, 5, 1>, 64, 64, 16, GettCfg<4, 4, 2, 2, 16, 5, 1>::NT, GettCfg<4, 4, 2, 2, 16, 5, 1>::SMEM, 1},
    // C1 "3M", 8 math warps of 32x16 (more registers per warp: double-buffered fragments)
    {k_gett<2, 4, 4, 2, 16, 5, 1>, 64, 64, 16, GettCfg<2, 4, 4, 2, 16, 5, 1>::NT, GettCfg<2, 4, 4, 2, 16, 5, 1>::SMEM, 1},
    // fused with the inner product that follows (FUSE = 1): variants of 6 (4M) and 7 (3M)
    {k_gett<4, 4, 4, 2, 16, 3, 0, 1>, 128, 64, 16, GettCfg<4, 4, 4, 2, 16, 3, 0, 1>::NT, GettCfg<4, 4, 4, 2, 16, 3, 0, 1>::SMEM, 1, GettCfg<4, 4, 4, 2, 16, 3, 0, 1>::DROWS},
    {k_gett<4, 4, 2, 2, 16, 5, 1, 1>, 64, 64, 16, GettCfg<4, 4, 2, 2, 16, 5, 1, 1>::NT, GettCfg<4, 4, 2, 2, 16, 5, 1, 1>::SMEM, 1, GettCfg<4, 4, 2, 2, 16, 5, 1, 1>::DROWS},
    // 3M with shared operand sums (gett3m.cuh, opt-in QTB_GETT_C1=5): 11 + 2 v plain, 12 + 2 v fused; v = QTB_G3: 0 = 5 ring stages, 1 = 4
#define G3(ST, SD, F) {k_gett3s<ST, SD, F>, 64, 64, 16, Gett3Cfg<ST, SD, F>::NT, Gett3Cfg<ST, SD, F>::SMEM, 1, Gett3Cfg<ST, SD, F>::DROWS}
    G3(5, 3, 0), G3(5, 3, 1), G3(4, 3, 0), G3(4, 3, 1),
#undef G3
};
static int fused_variant_of(int cfg) { return cfg == 6 ? 9 : cfg == 7 ? 10 : (cfg >= 11 && ((cfg - 11) & 1) == 0) ? cfg + 1 : -1; }
static int g3_variant() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_G3"); v = e ? std::max(0, std::min(1, atoi(e))) : 0; }
    return v;
}
static bool fusion_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_FUSE_DOT"); v = (e && !atoi(e)) ? 0 : 1; }
    return v == 1;
}
static int c1_variant() {
    static int v = -1;
    // 2 = "3M" complex product, 16 math warps, every warp adds its own operand sums (default);  5 = 3M with shared operand
    // sums (gett3m.cuh; correct, measured slower: profiles/r02_g3_shared_sums.txt);  3 = 3M, 8 math warps;  1 = 4M, 16 math
    // warps;  0 = 4M, 8 math warps
    if (v < 0) { const char *e = getenv("QTB_GETT_C1"); v = e ? atoi(e) : 2; }
    return v;
}

static int ilog2i(unsigned long long v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }

// ------------------------------------------------------------------------------------------------
// geometry: restates the leg bookkeeping of Network::ContractNodes (Network.h:739-769)
static int make_geom(int rA, int rB, int k, const int *posA, const int *posB, StepGeom &g) {
    if (rA < 0 || rB < 0 || rA > QTB_MAX_RANK || rB > QTB_MAX_RANK || k < 0 || k > rA || k > rB)
        return fail(QTB_ERR_INVALID, "bad ranks / shared-leg count");
    bool usedA[QTB_MAXR] = {false}, usedB[QTB_MAXR] = {false};
    for (int j = 0; j < k; j++) {
        const int a = posA[j], b = posB[j];
        if (a < 0 || a >= rA || b < 0 || b >= rB || usedA[a] || usedB[b] || (j > 0 && a <= posA[j - 1]))
            return fail(QTB_ERR_INVALID, "bad shared-leg map (pos_a must be strictly increasing, legs distinct)");
        usedA[a] = usedB[b] = true;
        g.posA[j] = a; g.posB[j] = b;
    }
    g.rA = rA; g.rB = rB; g.k = k; g.nfa = g.nfb = 0;
    for (int i = 0; i < rA; i++) if (!usedA[i]) g.freeA[g.nfa++] = i;
    for (int i = 0; i < rB; i++) if (!usedB[i]) g.freeB[g.nfb++] = i;
    g.rC = g.nfa + g.nfb;
    if (g.rC > QTB_MAX_RANK) return fail(QTB_ERR_INVALID, "result rank exceeds QTB_MAX_RANK");
    return QTB_OK;
}

static void make_devstep(const StepGeom &g, const double2 *A, const double2 *B, double2 *C, int kind, DevStep &st) {
    memset(&st, 0, sizeof(st));
    st.A = A; st.B = B; st.C = C;
    st.rA = g.rA; st.rB = g.rB; st.k = g.k; st.rC = g.rC; st.nfa = g.nfa; st.nfb = g.nfb; st.kind = kind;
    for (int i = 0; i < g.nfa; i++) st.shFree[i] = 2 * g.freeA[i];
    for (int i = 0; i < g.nfb; i++) st.shFree[g.nfa + i] = 2 * g.freeB[i];
    for (int i = 0; i < g.k; i++) {            // summed digit i <-> shared pair k-1-i (Network.h:912-916)
        st.shSumA[i] = 2 * g.posA[g.k - 1 - i];
        st.shSumB[i] = 2 * g.posB[g.k - 1 - i];
    }
}

static bool force_generic() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_FORCE_GENERIC"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}
static bool disable_micro() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_NO_MICRO"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}

static bool apply_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_NO_APPLY"); v = (e && atoi(e)) ? 0 : 1; }
    return v == 1;
}

static bool apply16_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_NO_APPLY16"); v = (e && atoi(e)) ? 0 : 1; }
    return v == 1;
}

static int choose_kind(const StepGeom &g, GettChoice &gc, int microLog4) {
    const unsigned long long U = g.units();
    if (!disable_micro() && U <= (1ull << (2 * microLog4)) && g.rA <= MICRO_MAX_RANK && g.rB <= MICRO_MAX_RANK && g.rC <= MICRO_MAX_RANK)
        return KIND_MICRO;
    const int bigFree = std::max(g.nfa, g.nfb), smallFree = std::min(g.nfa, g.nfb);
    if (!force_generic() && g.rC <= 2 && g.k >= 6) return KIND_REDUCE;       // long sums, <= 16 outputs: split-K
    // big tensor x tiny tensor (<= 64 elements, <= 4 outputs per free index): a pure stream, one thread per free index.
    // One exception stays with the tile kernel (0.87 vs 1.46 ms at rank 14): both shared legs are the big operand's two
    // lowest legs AND there are four outputs per free index.
    // ... and one more member: a rank-3 tensor sharing ONE leg (16 outputs per free index, 4 loads): write-dominated, the
    // tile kernel is issue-bound on it
    // (only when the big operand's legs come first in C.  With the small operand's legs first a thread's 16 outputs are a
    // 256-byte run, lanes 256 bytes apart: 1.27 ms at rank 13; sixteen lanes per free index with broadcast loads: 1.5-2.4 ms;
    // the tile kernel with its transposed 256-bit stores: 0.96 ms -- that case stays there)
    if (!force_generic() && apply_enabled() && bigFree >= 4 && g.k <= 2 &&
        (smallFree <= 1 || (smallFree == 2 && g.k == 1 && g.nfa > g.nfb && apply16_enabled()))) {
        const bool sw = g.nfb > g.nfa;
        const int *posX = sw ? g.posB : g.posA;
        const bool lowPair = g.k == 2 && std::min(posX[0], posX[1]) == 0 && std::max(posX[0], posX[1]) == 1;
        if (!(lowPair && smallFree == 1)) { gc.swap = sw; return KIND_APPLY; }
    }
    if (!force_generic() && bigFree >= 4 && g.k >= 1) {
        gc.swap = g.nfb > g.nfa;
        const int tk4 = (g.k == 1) ? 1 : 0;
        if (smallFree >= 3) gc.cfg = (tk4 == 0 && c1_variant() == 3) ? 8 : (tk4 == 0 && c1_variant() == 5) ? 11 + 2 * g3_variant() : (tk4 == 0 && c1_variant() == 2) ? 7 :
                                     (tk4 == 0 && c1_variant() == 1) ? 6 : 0 + tk4;
        else if (smallFree == 2) gc.cfg = 2 + tk4;
        else gc.cfg = 4 + tk4;
        return KIND_GETT;
    }
    return (g.rC >= 6) ? KIND_THREAD : KIND_WARP;
}

// xOrder (optional): which free leg of X each x digit stands for (a permutation of 0..nfx-1).  The default, ascending,
// makes C's leading legs the tile; the fused variant writes no C and picks the legs that suit its second operand.
static void build_gett(const StepGeom &g, const GettChoice &gc, const double2 *A, const double2 *B, double2 *C, GettParams &p,
                       const int *xOrder = nullptr) {
    const GettInst &inst = g_gett[gc.cfg];
    const int TMB = ilog2i(inst.TM), TNB = ilog2i(inst.TN), TKB = ilog2i(inst.TK);
    memset(&p, 0, sizeof(p));
    const bool sw = gc.swap;
    p.X = sw ? B : A; p.Y = sw ? A : B; p.C = C;
    const int nfx = sw ? g.nfb : g.nfa, nfy = sw ? g.nfa : g.nfb;
    const int *freeX = sw ? g.freeB : g.freeA, *freeY = sw ? g.freeA : g.freeB;
    const int cx0 = sw ? g.nfa : 0, cy0 = sw ? 0 : g.nfa;      // first C digit of the x / y legs
    p.xbits = 2 * nfx; p.ybits = 2 * nfy; p.kbits = 2 * g.k;
    for (int i = 0; i < nfx; i++) {
        const int f = xOrder ? xOrder[i] : i;
        for (int b = 0; b < 2; b++) { p.shXx[2 * i + b] = 2 * freeX[f] + b; p.shCx[2 * i + b] = 2 * (cx0 + f) + b; }
    }
    for (int i = 0; i < nfy; i++) for (int b = 0; b < 2; b++) { p.shYy[2 * i + b] = 2 * freeY[i] + b; p.shCy[2 * i + b] = 2 * (cy0 + i) + b; }
    // summation order is free: put the shared legs that sit lowest in either operand inside the k-chunk
    int ord[QTB_MAXR];
    for (int j = 0; j < g.k; j++) ord[j] = j;
    const int *posX = sw ? g.posB : g.posA, *posY = sw ? g.posA : g.posB;
    std::stable_sort(ord, ord + g.k, [&](int a, int b) { return std::min(posX[a], posY[a]) < std::min(posX[b], posY[b]); });
    for (int j = 0; j < g.k; j++) for (int b = 0; b < 2; b++) { p.shXk[2 * j + b] = 2 * posX[ord[j]] + b; p.shYk[2 * j + b] = 2 * posY[ord[j]] + b; }
    const unsigned long long Mx = 1ull << p.xbits, Ny = 1ull << p.ybits, K = 1ull << p.kbits;
    p.nyValid = (uint32_t)std::min<unsigned long long>(Ny, inst.TN);
    p.nyBits = ilog2i(p.nyValid);
    p.nTilesX = (uint32_t)(Mx / inst.TM);
    p.nTilesY = (uint32_t)std::max<unsigned long long>(1, Ny / inst.TN);
    p.nChunks = (uint32_t)(K / inst.TK);
    // load-order permutations: slot-id bits follow the operand's memory significance
    struct CB { int shift, coord; };
    std::vector<CB> v;
    for (int j = 0; j < TMB; j++) v.push_back({p.shXx[j], j});
    for (int j = 0; j < TKB; j++) v.push_back({p.shXk[j], TMB + j});
    std::sort(v.begin(), v.end(), [](const CB &a, const CB &b) { return a.shift < b.shift; });
    for (size_t j = 0; j < v.size(); j++) p.permX[j] = v[j].coord;
    v.clear();
    for (int j = 0; j < p.nyBits; j++) v.push_back({p.shYy[j], j});
    for (int j = 0; j < TKB; j++) v.push_back({p.shYk[j], TNB + j});
    std::sort(v.begin(), v.end(), [](const CB &a, const CB &b) { return a.shift < b.shift; });
    for (size_t j = 0; j < v.size(); j++) p.permY[j] = v[j].coord;
    // shared-memory layouts: contiguous-in-HBM dimension contiguous in shared memory (see GettParams)
    const int LDK = (inst.TK % 8 == 0) ? inst.TK + 4 : inst.TK;
    const bool xKFast = p.permX[0] >= TMB, yKFast = (p.nyBits + TKB > 0) && p.permY[0] >= TNB;
    if (xKFast) { p.xsX = (uint16_t)LDK; p.xsK = 1; } else { p.xsX = 1; p.xsK = (uint16_t)(inst.TM + 2); }
    if (yKFast) { p.ysY = (uint16_t)LDK; p.ysK = 1; } else { p.ysY = 1; p.ysK = (uint16_t)(inst.TN + 2); }
}

static void build_reduce(const StepGeom &g, const double2 *A, const double2 *B, double2 *C, double2 *partial, ReduceParams &p) {
    memset(&p, 0, sizeof(p));
    p.A = A; p.B = B; p.C = C; p.partial = partial;
    p.kbits = 2 * g.k;
    p.nTiles = (uint32_t)((1ull << p.kbits) / 256);
    // tile legs: the two lowest shared legs of A, then the two lowest of B (distinct pairs), so both
    // operands are read in 256-byte runs; the rest follow in A order
    std::vector<int> order;
    auto has = [&](int j) { return std::find(order.begin(), order.end(), j) != order.end(); };
    std::vector<int> byB(g.k);
    for (int j = 0; j < g.k; j++) byB[j] = j;
    std::sort(byB.begin(), byB.end(), [&](int a, int b) { return g.posB[a] < g.posB[b]; });
    order.push_back(0); order.push_back(1);                           // posA is increasing: pairs 0,1 are lowest in A
    for (int j = 0; j < g.k && order.size() < 4; j++) if (!has(byB[j])) order.push_back(byB[j]);
    for (int j = 0; j < g.k; j++) if (!has(j)) order.push_back(j);
    for (int j = 0; j < g.k; j++) for (int b = 0; b < 2; b++) { p.shA[2 * j + b] = 2 * g.posA[order[j]] + b; p.shB[2 * j + b] = 2 * g.posB[order[j]] + b; }
    const int NC = 1 << (2 * g.rC);
    for (int c = 0; c < NC; c++) {
        uint32_t oa = 0, ob = 0;
        for (int i = 0; i < g.nfa; i++) oa += (uint32_t)((c >> (2 * i)) & 3) << (2 * g.freeA[i]);
        for (int i = 0; i < g.nfb; i++) ob += (uint32_t)((c >> (2 * (g.nfa + i))) & 3) << (2 * g.freeB[i]);
        p.fA[c] = oa; p.fB[c] = ob;
    }
}

// A DMMA step whose result T is immediately contracted with another tensor D over ALL of T's legs (an inner product,
// rC = 0) can run fused: the tile kernel multiplies its accumulators with the matching D elements instead of storing T.
// `tIsA`: T is operand A of the inner product.  Returns false if the pair does not qualify.
static bool fusable_pair(const StepGeom &g1, int kind1, const GettChoice &gc1, const StepGeom &g2, bool tIsA) {
    if (!fusion_enabled() || kind1 != KIND_GETT || fused_variant_of(gc1.cfg) < 0) return false;
    if (g2.rC != 0 || g2.k != g1.rC || g2.rA != g2.rB || g2.rA != g1.rC) return false;
    (void)tIsA;
    return true;
}
// shDx / shDy of the fused kernel: where each x / y bit of the tile kernel lands inside D's element index
static void add_fusion(const StepGeom &g2, bool tIsA, const double2 *D, double2 *partial, const GettInst &inst, GettParams &p) {
    // leg l of T pairs with leg dLeg[l] of D
    int dLeg[QTB_MAXR];
    for (int j = 0; j < g2.k; j++) {
        if (tIsA) dLeg[g2.posA[j]] = g2.posB[j];
        else dLeg[g2.posB[j]] = g2.posA[j];
    }
    for (int j = 0; j < p.xbits; j++) p.shDx[j] = (uint8_t)(2 * dLeg[p.shCx[j] / 2] + (p.shCx[j] & 1));
    for (int j = 0; j < p.ybits; j++) p.shDy[j] = (uint8_t)(2 * dLeg[p.shCy[j] / 2] + (p.shCy[j] & 1));
    p.dotD = D;
    p.dotPartial = partial;
    // load order of one dotD piece (TM x drows): slot bits follow dotD's memory significance
    const int TMB = ilog2i(inst.TM), RB = ilog2i(inst.drows);
    struct CB { int shift, coord; };
    std::vector<CB> v;
    for (int j = 0; j < TMB; j++) v.push_back({p.shDx[j], j});
    for (int j = 0; j < RB; j++) v.push_back({p.shDy[j], TMB + j});
    std::sort(v.begin(), v.end(), [](const CB &a, const CB &b) { return a.shift < b.shift; });
    for (size_t j = 0; j < v.size(); j++) p.permD[j] = (uint8_t)v[j].coord;
}

// ------------------------------------------------------------------------------------------------
// TMA-fed tile kernel (gett_tma.cuh): eligibility, role assignment and tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)sym;
        else cudaGetLastError();
    }
    return fn;
}
// Opt-in (QTB_TMA=1): measured on B200 the TMA-fed variant is correct but 7 % (tile = one contiguous 16 KB run) to 19 %
// (tile legs picked for eligibility, 64-byte pieces) SLOWER than the cp.async gather on the config-2 rank-14 steps
// (profiles/r02_tma.txt), so the default path stays the cp.async ring.
static bool tma_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_TMA"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}


// the TMA instantiation: 3M complex product, 64x64x16 tiles, 16 math warps of 16x16, 6 ring stages of 32 KB
typedef GettTmaCfg<4, 4, 2, 2, 16, 6, 1> TmaCfg;
static void (*const g_gettTmaFn)(const GettTmaParams) = k_gett_tma<4, 4, 2, 2, 16, 6, 1>;

struct TileLeg { int pos; int role; int logical; };         // role 0: free (x or y), 1: shared (k); logical: index among the tile's legs of that role

// dense-index bit -> its contribution to the 16-byte bank group of a 64-byte-swizzled slot (see gett_tma.cuh):
// slot = idx ^ ((idx >> 3) & 3)
static int swizzle_vec(int p) { return p < 3 ? (1 << p) : (p < 5 ? (1 << (p - 3)) : 0); }
static bool independent3(int a, int b, int c) {
    if (!a || !b || !c) return false;
    return a != b && a != c && b != c && (a ^ b) != c;
}

// the tile must hold the operand's leg 0 (it is the contiguous inner dimension of the box); the other legs form runs of
// memory-adjacent legs -> tensor-map dimensions (at most four legs each); dense-index bit of every tile bit
struct OperandTile {
    std::vector<TileLeg> legs;                    // ascending position
    std::vector<std::pair<int, int>> dims;        // (first leg position, number of legs <= 4)
    int bitOf(int role, int logical, int b) const {
        for (size_t j = 0; j < legs.size(); j++) if (legs[j].role == role && legs[j].logical == logical) return 2 * (int)j + b;
        return -1;
    }
};
static bool analyse_tile(std::vector<TileLeg> legs, OperandTile &out) {
    std::sort(legs.begin(), legs.end(), [](const TileLeg &a, const TileLeg &b) { return a.pos < b.pos; });
    out.legs = legs;
    out.dims.clear();
    if (legs.empty() || legs[0].pos != 0) return false;
    for (size_t j = 1; j < legs.size();) {
        size_t e = j + 1;
        while (e < legs.size() && legs[e].pos == legs[e - 1].pos + 1) e++;
        for (size_t c = j; c < e; c += 4) out.dims.push_back({legs[c].pos, (int)std::min<size_t>(4, e - c)});
        j = e;
    }
    return out.dims.size() <= 3;
}

static int encode_tile_map(CUtensorMap *tm, const double2 *base, int rank, const OperandTile &t) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return fail(QTB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
    // dim0 = leg 0 x (re, im): 8 doubles = one 64-byte swizzle row; dim1 = the "offset" dimension in units of 4 elements
    cuuint64_t gdim[5] = {8, (cuuint64_t)1 << (2 * (rank - 1)), 1, 1, 1};
    cuuint64_t gstr[4] = {64, 0, 0, 0};
    cuuint32_t box[5] = {8, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    for (int d = 0; d < 3; d++) {
        if (d < (int)t.dims.size()) {
            gdim[2 + d] = (cuuint64_t)1 << (2 * t.dims[d].second);
            gstr[1 + d] = (cuuint64_t)16 << (2 * t.dims[d].first);
            box[2 + d] = (cuuint32_t)gdim[2 + d];
        } else {
            gstr[1 + d] = (cuuint64_t)16 << (2 * rank);       // unused dimension: size 1 beyond the tensor
        }
    }
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double2 *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QTB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return QTB_OK;
}

// Tries to set up the TMA-fed variant for a compute-bound tile step.  Returns false when the step does not qualify (too few
// free / shared legs for full 64 x 64 x 16 tiles, more than three runs per operand tile, no conflict-free role assignment,
// tensors of more than 2^31 elements): the caller then uses the cp.async gather of gett.cuh.
static bool build_gett_tma(const StepGeom &g, bool sw, const double2 *A, const double2 *B, double2 *C, GettTmaParams &P) {
    if (!tma_enabled() || !tma_encoder()) return false;
    const int TMB = 6, TNB = 6, TKB = 4;                                  // 64 x 64 x 16
    const int nfx = sw ? g.nfb : g.nfa, nfy = sw ? g.nfa : g.nfb;
    const int rX = sw ? g.rB : g.rA, rY = sw ? g.rA : g.rB;
    if (nfx < 3 || nfy < 3 || g.k < 2 || rX > 15 || rY > 15) return false;
    const int *freeX = sw ? g.freeB : g.freeA, *freeY = sw ? g.freeA : g.freeB;
    const int *posX = sw ? g.posB : g.posA, *posY = sw ? g.posA : g.posB;
    const int cx0 = sw ? g.nfa : 0, cy0 = sw ? 0 : g.nfa;
    int ord[QTB_MAXR];
    for (int j = 0; j < g.k; j++) ord[j] = j;
    std::stable_sort(ord, ord + g.k, [&](int a, int b) { return std::min(posX[a], posY[a]) < std::min(posX[b], posY[b]); });
    // candidate leg triples of an operand: its three lowest free legs, then every triple of its ten lowest that keeps the lowest
    auto candidates = [](int nf) {
        std::vector<std::array<int, 3>> out;
        out.push_back({0, 1, 2});
        const int m = std::min(nf, 10);
        for (int b = 1; b < m; b++) for (int c = b + 1; c < m; c++) if (!(b == 1 && c == 2)) out.push_back({0, b, c});
        return out;
    };
    for (const auto &xc : candidates(nfx)) {
        for (const auto &yc : candidates(nfy)) {
            OperandTile tx, ty;
            std::vector<TileLeg> lx, ly;
            for (int i = 0; i < 3; i++) { lx.push_back({freeX[xc[i]], 0, i}); ly.push_back({freeY[yc[i]], 0, i}); }
            for (int l = 0; l < 2; l++) { lx.push_back({posX[ord[l]], 1, l}); ly.push_back({posY[ord[l]], 1, l}); }
            if (!analyse_tile(lx, tx) || !analyse_tile(ly, ty)) continue;
            // roles: two k bits for (t0, t1) -- the same pair in both operands -- and one free bit per operand for g0
            int t0l = -1, t0b = 0, t1l = 0, t1b = 0, gxl = 0, gxb = 0, gyl = 0, gyb = 0;
            for (int a = 0; a < 4 && t0l < 0; a++) for (int b = 0; b < 4 && t0l < 0; b++) {
                if (a == b) continue;
                const int vx0 = swizzle_vec(tx.bitOf(1, a / 2, a % 2)), vx1 = swizzle_vec(tx.bitOf(1, b / 2, b % 2));
                const int vy0 = swizzle_vec(ty.bitOf(1, a / 2, a % 2)), vy1 = swizzle_vec(ty.bitOf(1, b / 2, b % 2));
                int fx = -1, fy = -1;
                for (int c = 0; c < 6 && fx < 0; c++) if (independent3(swizzle_vec(tx.bitOf(0, c / 2, c % 2)), vx0, vx1)) fx = c;
                for (int c = 0; c < 6 && fy < 0; c++) if (independent3(swizzle_vec(ty.bitOf(0, c / 2, c % 2)), vy0, vy1)) fy = c;
                if (fx >= 0 && fy >= 0) { t0l = a / 2; t0b = a % 2; t1l = b / 2; t1b = b % 2; gxl = fx / 2; gxb = fx % 2; gyl = fy / 2; gyb = fy % 2; }
            }
            if (t0l < 0) continue;
            // ---- fill the parameters
            memset(&P, 0, sizeof(P));
            GettParams &p = P.g;
            p.X = sw ? B : A; p.Y = sw ? A : B; p.C = C;
            p.xbits = 2 * nfx; p.ybits = 2 * nfy; p.kbits = 2 * g.k;
            p.nyValid = 64; p.nyBits = 6;
            // logical k bits: t0, t1, then the chunk's other two bits, then the remaining shared legs (k-chunk index)
            struct KB { int l, b; };
            std::vector<KB> kb = {{t0l, t0b}, {t1l, t1b}};
            for (int a = 0; a < 4; a++) if (!((a / 2 == t0l && a % 2 == t0b) || (a / 2 == t1l && a % 2 == t1b))) kb.push_back({a / 2, a % 2});
            for (int j = 0; j < TKB; j++) {
                p.shXk[j] = (uint8_t)(2 * posX[ord[kb[j].l]] + kb[j].b); p.shYk[j] = (uint8_t)(2 * posY[ord[kb[j].l]] + kb[j].b);
                P.ikX[j] = (uint8_t)tx.bitOf(1, kb[j].l, kb[j].b); P.ikY[j] = (uint8_t)ty.bitOf(1, kb[j].l, kb[j].b);
            }
            for (int j = 2; j < g.k; j++) for (int b = 0; b < 2; b++) { p.shXk[2 * j + b] = (uint8_t)(2 * posX[ord[j]] + b); p.shYk[2 * j + b] = (uint8_t)(2 * posY[ord[j]] + b); }
            // logical x bits: g0 first, then the tile's other five bits by C significance, then the remaining free legs (tile index)
            auto fillFree = [&](const std::array<int, 3> &cand, const int *freeLegs, int nf, int c0, int g0l, int g0b, const OperandTile &t,
                                uint8_t *shOp, uint8_t *shC, uint8_t *idxPos) {
                struct FB { int shOp, shC, idx; };
                std::vector<FB> bits;
                bits.push_back({2 * freeLegs[cand[g0l]] + g0b, 2 * (c0 + cand[g0l]) + g0b, t.bitOf(0, g0l, g0b)});
                std::vector<FB> rest;
                for (int l = 0; l < 3; l++) for (int b = 0; b < 2; b++) if (!(l == g0l && b == g0b)) rest.push_back({2 * freeLegs[cand[l]] + b, 2 * (c0 + cand[l]) + b, t.bitOf(0, l, b)});
                std::sort(rest.begin(), rest.end(), [](const FB &a, const FB &b) { return a.shC < b.shC; });
                bits.insert(bits.end(), rest.begin(), rest.end());
                for (int j = 0; j < 6; j++) { shOp[j] = (uint8_t)bits[j].shOp; shC[j] = (uint8_t)bits[j].shC; idxPos[j] = (uint8_t)bits[j].idx; }
                int n = 6;
                for (int f = 0; f < nf; f++) {
                    if (f == cand[0] || f == cand[1] || f == cand[2]) continue;
                    for (int b = 0; b < 2; b++) { shOp[n] = (uint8_t)(2 * freeLegs[f] + b); shC[n] = (uint8_t)(2 * (c0 + f) + b); n++; }
                }
            };
            fillFree(xc, freeX, nfx, cx0, gxl, gxb, tx, p.shXx, p.shCx, P.ixX);
            fillFree(yc, freeY, nfy, cy0, gyl, gyb, ty, p.shYy, p.shCy, P.iyY);
            const unsigned long long Mx = 1ull << p.xbits, Ny = 1ull << p.ybits, K = 1ull << p.kbits;
            p.nTilesX = (uint32_t)(Mx / 64); p.nTilesY = (uint32_t)(Ny / 64); p.nChunks = (uint32_t)(K / 16);
            auto swz = [](uint32_t idx) { return (uint16_t)((idx ^ ((idx >> 3) & 3u)) << 4); };
            for (int i = 0; i < 4; i++) {            // x (y) advances by 8 per fragment index: logical bits 3, 4
                uint32_t ix = 0, iy = 0;
                for (int b = 0; b < 2; b++) if ((i >> b) & 1) { ix |= 1u << P.ixX[3 + b]; iy |= 1u << P.iyY[3 + b]; }
                P.cXi[i] = swz(ix); P.cYj[i] = swz(iy);
            }
            for (int kk = 0; kk < 4; kk++) {         // k advances by 4 per k-step: logical k bits 2, 3
                uint32_t ix = 0, iy = 0;
                for (int b = 0; b < 2; b++) if ((kk >> b) & 1) { ix |= 1u << P.ikX[2 + b]; iy |= 1u << P.ikY[2 + b]; }
                P.cXkk[kk] = swz(ix); P.cYkk[kk] = swz(iy);
            }
            if (encode_tile_map(&P.tmX, p.X, rX, tx) != QTB_OK || encode_tile_map(&P.tmY, p.Y, rY, ty) != QTB_OK) return false;
            (void)TMB; (void)TNB;
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// per-rank pooled device memory
struct Pool {
    std::vector<void *> freeList[QTB_MAX_RANK + 1];
    std::vector<void *> chunks;                 // cudaMalloc'd blocks owned by this pool
    uint8_t *slab = nullptr; size_t slabLeft = 0;
    long long reserved = 0, live = 0, peakLive = 0;
    static size_t bytes(int rank) { return (size_t)16 << (2 * rank); }
    static const size_t SLAB = (size_t)8 << 20;   // small tensors (<= 64 KB) are carved from 8 MB slabs

    int alloc(int rank, void **out) {
        std::vector<void *> &fl = freeList[rank];
        const size_t b = bytes(rank);
        if (!fl.empty()) { *out = fl.back(); fl.pop_back(); }
        else if (b <= (64u << 10)) {
            const size_t need = std::max<size_t>(b, 256);        // keep 256 B alignment
            if (slabLeft < need) {
                void *s = nullptr;
                CU(cudaMalloc(&s, SLAB));
                chunks.push_back(s); reserved += SLAB;
                slab = (uint8_t *)s; slabLeft = SLAB;
            }
            *out = slab; slab += need; slabLeft -= need;
        } else {
            void *p = nullptr;
            cudaError_t e = cudaMalloc(&p, b);
            if (e == cudaErrorMemoryAllocation) {                  // give cached blocks back and retry once
                cudaGetLastError();
                trim();
                e = cudaMalloc(&p, b);
            }
            if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? QTB_ERR_OOM : QTB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
            chunks.push_back(p); reserved += (long long)b;
            *out = p;
        }
        live += (long long)b; peakLive = std::max(peakLive, live);
        return QTB_OK;
    }
    void release(int rank, void *p) { freeList[rank].push_back(p); live -= (long long)bytes(rank); }
    void trim() {       // free cached big blocks (callers have synchronised the stream)
        cudaDeviceSynchronize();
        for (int r = 0; r <= QTB_MAX_RANK; r++) {
            if (bytes(r) <= (64u << 10)) continue;
            for (void *p : freeList[r]) {
                cudaFree(p); reserved -= (long long)bytes(r);
                chunks.erase(std::find(chunks.begin(), chunks.end(), p));
            }
            freeList[r].clear();
        }
    }
    void destroy() {
        for (void *p : chunks) cudaFree(p);
        chunks.clear();
        for (auto &f : freeList) f.clear();
        slab = nullptr; slabLeft = 0; reserved = live = 0;
    }
};

struct qtb_tensor_s {
    double2 *d = nullptr;
    int rank = 0;
    bool hasData = false;       // something has been uploaded / contracted into it (possibly still pending)
    bool pooled = false;        // d belongs to the ctx pool
};

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (the library must load on boxes without NCCL)
struct NcclId { char b[QTB_UNIQUE_ID_BYTES]; };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_ncclMu;
static int load_nccl() {
    std::lock_guard<std::mutex> lk(g_ncclMu);
    if (g_nccl.h) return QTB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(QTB_ERR_NCCL, "libnccl.so.2 not found");
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(QTB_ERR_NCCL, "NCCL symbols missing");
    g_nccl.h = h;
    return QTB_OK;
}

// ------------------------------------------------------------------------------------------------
struct PendingStep { DevStep st; uint32_t level; };
struct PendingUpload { double2 *dst; size_t payloadOff; uint32_t elems; };

struct TraceRec { cudaEvent_t e0, e1; int rA, rB, k, kernel; };

struct qtb_scalar_read_s { double *pinned = nullptr; cudaEvent_t done = nullptr; };

struct qtb_ctx_s {
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    Pool pool;
    // staging ring (pinned host mirror + device): micro-batch blobs and small uploads
    uint8_t *ringHost = nullptr, *ringDev = nullptr;
    size_t ringSize = (size_t)32 << 20, ringCur = 0;
    cudaEvent_t ringEvent = nullptr; bool ringEventValid = false;
    double *scalarPinned = nullptr;
    std::vector<qtb_scalar_read_s *> freeReads;   // recycled handles of qtb_read_scalar_begin/end
    uint64_t *zeroOffsetDev = nullptr;          // blobOffsets[0] = 0 for single-blob launches
    double2 *reduceScratch = nullptr;           // split-K partials: [REDUCE_MAX_BLOCKS][16]
    double2 *scratchOverride = nullptr;         // a plan's private partials while its launches are enqueued (plans may run side by side)
    double2 *partials() const { return scratchOverride ? scratchOverride : reduceScratch; }
    // fork/join streams for independent plans (qtb_plans_run_batched)
    std::vector<cudaStream_t> auxStreams; std::vector<cudaEvent_t> auxEvents; cudaEvent_t forkEvent = nullptr;
    // deferred micro work
    std::vector<PendingStep> pending;
    std::vector<PendingUpload> pendingUploads;
    std::vector<uint8_t> payload;
    std::unordered_map<const void *, uint32_t> producedLevel;     // tensor buffer -> level of the pending step writing it
    std::unordered_map<const void *, uint32_t> readLevel;         // tensor buffer -> highest level of a pending step READING it
    std::vector<std::pair<int, void *>> deferredFrees;            // (rank, ptr) released after the next flush
    // stats / trace
    qtb_stats stats{};
    bool trace = false;
    std::vector<TraceRec> traceRecs;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    int microLog4 = MICRO_DEFAULT_LOG4;
    // a big DMMA step that has been requested but not launched yet: if the very next step is the inner product of its
    // result with another tensor, both run as one fused kernel (the intermediate never touches HBM)
    struct Held { bool active = false; StepGeom g; GettChoice gc{0, false}; const double2 *A = nullptr, *B = nullptr; double2 *C = nullptr; } held;
    double *batchOut = nullptr; size_t batchOutCap = 0;      // pinned gather buffer of qtb_plans_run_batched
    // NCCL
    void *comm = nullptr; int nRanks = 1, rank = 0;
    double *commBuf = nullptr; size_t commBufElems = 0;
};

static int ensure_device(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    CU(cudaSetDevice(ctx->device));
    return QTB_OK;
}

// streaming step: X = the operand with the larger free dimension, Y = the tiny one (apply.cuh)
static int launch_apply(qtb_ctx *ctx, const StepGeom &g, bool swap, const double2 *A, const double2 *B, double2 *C, cudaStream_t s) {
    ApplyParams p;
    memset(&p, 0, sizeof(p));
    p.X = swap ? B : A; p.Y = swap ? A : B; p.C = C;
    const int nfx = swap ? g.nfb : g.nfa, nfy = swap ? g.nfa : g.nfb;
    const int *posX = swap ? g.posB : g.posA, *posY = swap ? g.posA : g.posB, *freeY = swap ? g.freeA : g.freeB;
    p.M = 1ull << (2 * nfx);
    p.nHoles = (uint8_t)g.k;
    p.yFirst = swap ? 1 : 0;                       // C legs = A-free then B-free (Network.h:810-812)
    int holes[2] = {0, 0};
    for (int j = 0; j < g.k; j++) holes[j] = 2 * posX[j];
    if (g.k == 2 && holes[0] > holes[1]) std::swap(holes[0], holes[1]);
    for (int j = 0; j < g.k; j++) p.holeBit[j] = (uint8_t)holes[j];
    const int K = 1 << (2 * g.k), N = 1 << (2 * nfy);
    for (int sv = 0; sv < K; sv++) {
        uint32_t ox = 0, oy = 0;
        for (int j = 0; j < g.k; j++) { const uint32_t d = (sv >> (2 * j)) & 3; ox += d << (2 * posX[j]); oy += d << (2 * posY[j]); }
        p.kOff[sv] = ox;
        for (int y = 0; y < N; y++) {
            uint32_t o = oy;
            for (int i = 0; i < nfy; i++) o += ((y >> (2 * i)) & 3) << (2 * freeY[i]);
            p.yIdx[sv][y] = (uint8_t)o;
        }
    }
    // both shared legs are X's two lowest legs: four lanes per output, each reading a 64-byte quarter of the 256-byte run
    if (g.k == 2 && std::min(posX[0], posX[1]) == 0 && std::max(posX[0], posX[1]) == 1) {
        uint8_t byMem[16][16];
        memset(byMem, 0, sizeof(byMem));
        for (int sv = 0; sv < K; sv++) memcpy(byMem[p.kOff[sv]], p.yIdx[sv], 16);     // kOff is a permutation of 0..15 here
        memcpy(p.yIdx, byMem, sizeof(byMem));
        const unsigned grid4 = (unsigned)(p.M * 4 / 256);
        if (nfy != 0) return fail(QTB_ERR_INVALID, "step outside the streaming class");
        k_apply_lowpair<1><<<grid4, 256, 0, s>>>(p);
        CU(cudaGetLastError());
        ctx->stats.launches++;
        return QTB_OK;
    }
    // the single shared leg is X's leg 0: a thread's four summands are one 64-byte run -> two 256-bit loads
    p.lowRun = (g.k == 1 && posX[0] == 0) ? 1 : 0;
    const unsigned grid = (unsigned)(p.M / 256);
    switch (g.k * 4 + nfy) {
        case 0: k_apply<1, 1><<<grid, 256, 0, s>>>(p); break;
        case 1: k_apply<1, 4><<<grid, 256, 0, s>>>(p); break;
        case 4: k_apply<4, 1><<<grid, 256, 0, s>>>(p); break;
        case 5: k_apply<4, 4><<<grid, 256, 0, s>>>(p); break;
        case 6: k_apply<4, 16><<<grid, 256, 0, s>>>(p); break;
        case 8: k_apply<16, 1><<<grid, 256, 0, s>>>(p); break;
        case 9: k_apply<16, 4><<<grid, 256, 0, s>>>(p); break;
        default: return fail(QTB_ERR_INVALID, "step outside the streaming class");
    }
    CU(cudaGetLastError());
    ctx->stats.launches++;
    return QTB_OK;
}

// QTB_PDL=1: k_gett launches carry the programmatic-dependent-launch attribute (gett.cuh: griddepcontrol), so the next tile-kernel
// launch builds its tables on the SMs the previous one has left.  Measured neutral on config 2 (14.589 vs 14.585 ms per term) and
// slightly negative with two plan lanes competing for SMs (sliced config-2 term 14.45 -> 14.67 ms), so it is off by default.
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_PDL"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}
static int launch_gett_kernel(const GettInst &inst, unsigned grid, const GettParams &p, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((unsigned)inst.NT); cfg.dynamicSmemBytes = inst.smem; cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, inst.fn, p));
    return QTB_OK;
}
static int launch_gett(qtb_ctx *ctx, const GettParams &p, int cfg, cudaStream_t s) {
    const GettInst &inst = g_gett[cfg];
    const unsigned nTiles = p.nTilesX * p.nTilesY;
    const unsigned grid = std::min<unsigned>(nTiles, (unsigned)(ctx->numSMs * inst.occ));
    ST(launch_gett_kernel(inst, grid, p, s));
    ctx->stats.launches++;
    return QTB_OK;
}

static const unsigned REDUCE_MAX_BLOCKS = 148 * 8;
static int launch_reduce(qtb_ctx *ctx, const StepGeom &g, const double2 *A, const double2 *B, double2 *C, cudaStream_t s) {
    ReduceParams p;
    build_reduce(g, A, B, C, ctx->partials(), p);
    if (g.rC == 0) {
        // inner product: each operand streamed in its own memory order, permutation undone in shared memory
        DotParams d;
        memset(&d, 0, sizeof(d));
        d.A = A; d.B = B; d.partial = ctx->partials(); d.nTiles = p.nTiles; d.kbits = p.kbits;
        memcpy(d.shA, p.shA, 32); memcpy(d.shB, p.shB, 32);
        int ia[8], ib[8];
        for (int j = 0; j < 8; j++) ia[j] = ib[j] = j;
        std::sort(ia, ia + 8, [&](int x, int y) { return p.shA[x] < p.shA[y]; });
        std::sort(ib, ib + 8, [&](int x, int y) { return p.shB[x] < p.shB[y]; });
        for (int j = 0; j < 8; j++) { d.permA[j] = (uint8_t)ia[j]; d.permB[j] = (uint8_t)ib[j]; }
        const unsigned grid = std::min<unsigned>((p.nTiles + QTB_DOT_T - 1) / QTB_DOT_T, std::min<unsigned>(REDUCE_MAX_BLOCKS, (unsigned)ctx->numSMs * 6));
        k_dot<<<grid, 256, 0, s>>>(d);
        k_reduce_final<1><<<1, 32, 0, s>>>(ctx->partials(), C, grid);
        CU(cudaGetLastError());
        ctx->stats.launches += 2;
        return QTB_OK;
    }
    const unsigned grid = std::min<unsigned>(p.nTiles, std::min<unsigned>(REDUCE_MAX_BLOCKS, (unsigned)ctx->numSMs * 8));
    switch (g.rC) {
        case 1: k_reduce<4><<<grid, 256, 0, s>>>(p); k_reduce_final<4><<<1, 128, 0, s>>>(p.partial, C, grid); break;
        default: k_reduce<16><<<grid, 256, 0, s>>>(p); k_reduce_final<16><<<1, 512, 0, s>>>(p.partial, C, grid); break;
    }
    CU(cudaGetLastError());
    ctx->stats.launches += 2;
    return QTB_OK;
}

static int launch_generic(qtb_ctx *ctx, const DevStep &st, cudaStream_t s) {
    const unsigned long long NC = 1ull << (2 * st.rC);
    if (st.kind == KIND_THREAD) {
        const unsigned grid = (unsigned)std::min<unsigned long long>((NC + 255) / 256, (unsigned long long)ctx->numSMs * 32);
        k_step_thread<<<grid, 256, 0, s>>>(st);
    } else {
        const unsigned grid = (unsigned)std::min<unsigned long long>((NC + 7) / 8, (unsigned long long)ctx->numSMs * 8);
        k_step_warp<<<grid, 256, 0, s>>>(st);
    }
    CU(cudaGetLastError());
    ctx->stats.launches++;
    return QTB_OK;
}

// ---- debugging aid: per-level clock stamps of compiled micro blobs (QTB_MICRO_TIMELINE=1; qtb_debug_dump_micro_timelines) -------
struct MicroTimeline { unsigned long long *dev = nullptr; uint32_t nLevels = 0; std::vector<uint32_t> items, maxSerial; };
static std::vector<MicroTimeline> g_microTimelines;
static bool micro_timeline_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_MICRO_TIMELINE"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}

static int launch_micro_group(qtb_ctx *ctx, const uint8_t *blobBase, const uint64_t *offsetDev, long long units, cudaStream_t s);

// Lane layout of one micro-step (k_micro): 2^lg lanes share an output along the summed index, P = 32 >> lg outputs per pass,
// 2^lp <= 4 passes per item; returns the serial multiply-add chain of one lane, 2^lp * max(1, K >> lg).  G is at least 32 / NC
// (tiny results still use the whole warp); passes shrink first and then G grows while the chain is longer than `target`.
static uint32_t micro_item_layout(uint32_t NC, uint32_t K, uint32_t target, uint32_t &lg, uint32_t &lp) {
    lg = 0; lp = 2;
    while ((32u >> lg) > NC) lg++;                                    // P = 32 / G <= NC
    while (lp > 0 && ((32u >> lg) << lp) > NC) lp--;                  // passes * P <= NC
    auto serialOf = [&](uint32_t l, uint32_t q) { return (1u << q) * std::max<uint32_t>(1u, K >> l); };
    while (serialOf(lg, lp) > target) {                               // fewer passes first (more items, same lanes per output)
        if (lp > 0) lp--;
        else if (lg < 5 && (2u << lg) <= K) lg++;
        else break;
    }
    return serialOf(lg, lp);
}

// ---- micro-batch blob assembly -----------------------------------------------------------------
// Builds: MicroHeader | levelItemStart | items | steps (copy steps first at level 0) | payload
static size_t build_micro_blob(const std::vector<PendingStep> &steps, const std::vector<PendingUpload> &ups,
                               const std::vector<uint8_t> &payload, std::vector<uint8_t> &blob, const uint8_t *devBase,
                               const void *prefetchPtr = nullptr, size_t prefetchBytes = 0) {
    // level 0 = upload copies and steps with no pending producer; a step's level is
    // 1 + max(level of the pending op producing each operand)
    uint32_t nLevels = ups.empty() ? 0 : 1;
    for (const auto &s : steps) nLevels = std::max(nLevels, s.level + 1);
    const uint32_t nSteps = (uint32_t)(steps.size() + ups.size());
    std::vector<std::vector<MicroItem>> perLevel(nLevels);
    std::vector<DevStep> all(nSteps);
    uint32_t idx = 0;
    for (const auto &u : ups) {
        DevStep &st = all[idx];
        memset(&st, 0, sizeof(st));
        st.kind = KIND_COPY; st.C = u.dst; st.A = nullptr;       // A patched below (needs payload offset)
        st.rC = 0; st.k = 0;
        // copy steps reuse the item machinery: chunk = block of QTB_MICRO_CHUNK elements
        const uint32_t nChunks = (u.elems + QTB_MICRO_CHUNK - 1) / QTB_MICRO_CHUNK;
        for (uint32_t c = 0; c < nChunks; c++) perLevel[0].push_back({idx, c});
        idx++;
    }
    // Lane layout of a step (k_micro): G = 2^lg lanes share one output along the summed index, 32 / G outputs per pass, 2^lp <= 4
    // passes per item.  G is at least 32 / NC (tiny results still use the whole warp) and grows while one item's serial
    // multiply-add chain is longer than a fair share of its level: a level lasts as long as its slowest warp, and a step with
    // few outputs and a long sum (a QAOA p=2 term has (NC, K) = (256, 256) steps alone in their level) would otherwise keep
    // 2 of the CTA's 32 warps busy for 1024 dependent FMAs per lane.
    std::vector<double> levelWork(nLevels, 0.0);
    for (const auto &s : steps) levelWork[s.level] += (double)(1ull << (2 * (s.st.rC + s.st.k))) / 32.0;
    struct Heavy { uint32_t serial; MicroItem it; };
    std::vector<std::vector<Heavy>> sorted(nLevels);
    for (const auto &s : steps) {
        all[idx] = s.st;
        const uint32_t NC = 1u << (2 * s.st.rC), K = 1u << (2 * s.st.k);
        uint32_t lg, lp;
        const uint32_t target = (uint32_t)std::max(16.0, levelWork[s.level] / (QTB_MICRO_THREADS / 32));
        const uint32_t itemSerial = micro_item_layout(NC, K, target, lg, lp);
        auto serialOf = [&](uint32_t, uint32_t) { return itemSerial; };
        const uint32_t perItem = (32u >> lg) << lp;                       // outputs of one item
        const uint32_t nChunks = NC / perItem;
        for (uint32_t c = 0; c < nChunks; c++) sorted[s.level].push_back({serialOf(lg, lp), {idx, c | (lg << 24) | (lp << 29)}});
        idx++;
    }
    for (uint32_t l = 0; l < nLevels; l++) {                              // heaviest items first: warps take items round-robin
        std::stable_sort(sorted[l].begin(), sorted[l].end(), [](const Heavy &a, const Heavy &b) { return a.serial > b.serial; });
        for (const auto &h : sorted[l]) perLevel[l].push_back(h.it);
    }
    uint32_t nItems = 0;
    for (auto &v : perLevel) nItems += (uint32_t)v.size();
    size_t off = sizeof(MicroHeader) + sizeof(uint32_t) * (nLevels + 1) + sizeof(MicroItem) * nItems;
    off = (off + 15) & ~(size_t)15;
    const size_t stepsOff = off;
    off += sizeof(DevStep) * nSteps;
    off = (off + 15) & ~(size_t)15;
    const size_t payloadOff = off;
    off += payload.size();
    blob.assign(off, 0);
    MicroHeader hdr{nLevels, nItems, nSteps, (uint32_t)stepsOff, (uint32_t)payloadOff, 0u, 0ull, 0ull};
    if (micro_timeline_enabled() && devBase == nullptr && ups.empty()) {          // compiled plans only (their blobs live as long as the plan)
        MicroTimeline tl;
        tl.nLevels = nLevels;
        if (cudaMalloc((void **)&tl.dev, (size_t)(nLevels + 2) * 8) == cudaSuccess) {
            cudaMemset(tl.dev, 0, (size_t)(nLevels + 2) * 8);
            for (uint32_t l = 0; l < nLevels; l++) {
                uint32_t maxSerial = 0;
                for (const auto &h : sorted[l]) maxSerial = std::max(maxSerial, h.serial);
                tl.items.push_back((uint32_t)perLevel[l].size()); tl.maxSerial.push_back(maxSerial);
            }
            hdr.timelinePtr = (uint64_t)tl.dev;
            g_microTimelines.push_back(tl);
        } else cudaGetLastError();
    }
    if (prefetchPtr && prefetchBytes) { hdr.prefetchPtr = (uint64_t)prefetchPtr; hdr.prefetchBytes = (uint32_t)std::min<size_t>(prefetchBytes, 8u << 20); }
    else if (!payload.empty() && devBase) { hdr.prefetchPtr = (uint64_t)(devBase + payloadOff); hdr.prefetchBytes = (uint32_t)payload.size(); }
    memcpy(blob.data(), &hdr, sizeof(hdr));
    uint32_t *lis = reinterpret_cast<uint32_t *>(blob.data() + sizeof(MicroHeader));
    MicroItem *items = reinterpret_cast<MicroItem *>(lis + nLevels + 1);
    uint32_t cur = 0;
    for (uint32_t l = 0; l < nLevels; l++) {
        lis[l] = cur;
        for (const auto &it : perLevel[l]) items[cur++] = it;
    }
    lis[nLevels] = cur;
    // patch copy sources (device address of the payload inside the device-side blob)
    for (size_t i = 0; i < ups.size(); i++) {
        all[i].A = reinterpret_cast<const double2 *>(devBase + payloadOff + ups[i].payloadOff);
        // element count travels in the (otherwise unused) B pointer slot
        all[i].B = reinterpret_cast<const double2 *>((uintptr_t)ups[i].elems);
    }
    memcpy(blob.data() + stepsOff, all.data(), sizeof(DevStep) * nSteps);
    if (!payload.empty()) memcpy(blob.data() + payloadOff, payload.data(), payload.size());
    return off;
}


static int ring_reserve(qtb_ctx *ctx, size_t bytes, size_t &off) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > ctx->ringSize) return fail(QTB_ERR_INVALID, "micro batch larger than the staging ring");
    if (ctx->ringCur + bytes > ctx->ringSize) {
        // wrap: everything enqueued so far must have consumed its staging bytes
        if (ctx->ringEventValid) CU(cudaEventSynchronize(ctx->ringEvent));
        ctx->ringCur = 0;
    }
    off = ctx->ringCur;
    ctx->ringCur += bytes;
    return QTB_OK;
}

static int enqueue_big(qtb_ctx *ctx, const StepGeom &g, int kind, const GettChoice &gc, const double2 *A, const double2 *B,
                       double2 *C, cudaStream_t s);

static int launch_held_locked(qtb_ctx *ctx) {
    if (!ctx->held.active) return QTB_OK;
    ctx->held.active = false;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->trace) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, ctx->stream)); }
    ST(enqueue_big(ctx, ctx->held.g, KIND_GETT, ctx->held.gc, ctx->held.A, ctx->held.B, ctx->held.C, ctx->stream));
    if (ctx->trace) { CU(cudaEventRecord(e1, ctx->stream)); ctx->traceRecs.push_back({e0, e1, ctx->held.g.rA, ctx->held.g.rB, ctx->held.g.k, KIND_GETT}); }
    return QTB_OK;
}

static int flush_locked(qtb_ctx *ctx) {
    ST(launch_held_locked(ctx));          // program order: the held step was requested before anything still pending
    if (ctx->pending.empty() && ctx->pendingUploads.empty()) {
        for (auto &f : ctx->deferredFrees) ctx->pool.release(f.first, f.second);
        ctx->deferredFrees.clear();
        return QTB_OK;
    }
    ST(ensure_device(ctx));
    std::vector<uint8_t> blob;
    // two passes: size first (device base unknown until the ring slot is known)
    size_t need = build_micro_blob(ctx->pending, ctx->pendingUploads, ctx->payload, blob, nullptr);
    size_t off = 0;
    ST(ring_reserve(ctx, need, off));
    build_micro_blob(ctx->pending, ctx->pendingUploads, ctx->payload, blob, ctx->ringDev + off);
    memcpy(ctx->ringHost + off, blob.data(), blob.size());
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->trace) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, ctx->stream)); }
    CU(cudaMemcpyAsync(ctx->ringDev + off, ctx->ringHost + off, blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.bytes_h2d += (long long)blob.size();
    long long groupUnits = 0;
    for (const auto &ps : ctx->pending) groupUnits += 1ll << (2 * (ps.st.rC + ps.st.k));
    ST(launch_micro_group(ctx, ctx->ringDev + off, ctx->zeroOffsetDev, groupUnits, ctx->stream));
    CU(cudaEventRecord(ctx->ringEvent, ctx->stream));
    ctx->ringEventValid = true;
    if (ctx->trace) { CU(cudaEventRecord(e1, ctx->stream)); ctx->traceRecs.push_back({e0, e1, 0, 0, (int)ctx->pending.size(), KIND_MICRO}); }
    ctx->stats.launches++;
    ctx->stats.micro_steps += (long long)ctx->pending.size();
    ctx->pending.clear(); ctx->pendingUploads.clear(); ctx->payload.clear(); ctx->producedLevel.clear(); ctx->readLevel.clear();
    for (auto &f : ctx->deferredFrees) ctx->pool.release(f.first, f.second);
    ctx->deferredFrees.clear();
    return QTB_OK;
}

static int ensure_buffer(qtb_ctx *ctx, qtb_tensor t) {
    if (t->d) return QTB_OK;
    void *p = nullptr;
    ST(ctx->pool.alloc(t->rank, &p));
    t->d = (double2 *)p; t->pooled = true;
    return QTB_OK;
}

// fused DMMA step + inner product: T = contract(A1, B1) is never materialised; out = <T, D>
static int enqueue_fused(qtb_ctx *ctx, const StepGeom &g1, const GettChoice &gc1, const double2 *A1, const double2 *B1,
                         const StepGeom &g2, bool tIsA, const double2 *D, double2 *out, cudaStream_t s) {
    GettParams p;
    // Tile legs of the fused step.  No C is written, so the x digits inside a tile are free to choose: keep X's lowest
    // free leg (its gather keeps >= 64-byte runs) and add the legs whose partners are D's LOWEST legs, so that the matching
    // D tile is made of whole sectors (with C's leading legs it was 16 useful bytes per 32-byte sector: 8.6 GB for 4.3).
    int xOrder[QTB_MAXR];
    {
        const bool sw = gc1.swap;
        const int nfx = sw ? g1.nfb : g1.nfa, nfy = sw ? g1.nfa : g1.nfb, cx0 = sw ? g1.nfa : 0, cy0 = sw ? 0 : g1.nfa;
        int dLeg[QTB_MAXR];                                 // leg l of T pairs with leg dLeg[l] of D
        for (int j = 0; j < g2.k; j++) { if (tIsA) dLeg[g2.posA[j]] = g2.posB[j]; else dLeg[g2.posB[j]] = g2.posA[j]; }
        const int tileLegs = ilog2i(g_gett[fused_variant_of(gc1.cfg)].TM) / 2;
        bool dCovered[QTB_MAXR] = {false}, taken[QTB_MAXR] = {false};
        for (int i = 0; i < std::min(nfy, ilog2i(g_gett[fused_variant_of(gc1.cfg)].TN) / 2); i++) dCovered[dLeg[cy0 + i]] = true;   // the tile's y legs
        int n = 0;
        xOrder[n++] = 0; taken[0] = true; dCovered[dLeg[cx0]] = true;
        for (int d = 0; d < g2.k && n < std::min(tileLegs, nfx); d++) {      // D's legs from the bottom up
            if (dCovered[d]) continue;
            for (int i = 0; i < nfx; i++)
                if (!taken[i] && dLeg[cx0 + i] == d) { xOrder[n++] = i; taken[i] = true; dCovered[d] = true; break; }
        }
        for (int i = 0; i < nfx; i++) if (!taken[i]) xOrder[n++] = i;
    }
    build_gett(g1, gc1, A1, B1, nullptr, p, xOrder);
    const int cfg = fused_variant_of(gc1.cfg);
    const GettInst &inst = g_gett[cfg];
    add_fusion(g2, tIsA, D, ctx->partials(), inst, p);
    const unsigned nTiles = p.nTilesX * p.nTilesY;
    const unsigned grid = std::min<unsigned>(nTiles, (unsigned)(ctx->numSMs * inst.occ));
    ST(launch_gett_kernel(inst, grid, p, s));
    k_reduce_final<1><<<1, 32, 0, s>>>(ctx->partials(), out, grid);
    CU(cudaGetLastError());
    ctx->stats.launches += 2;
    return QTB_OK;
}

// enqueue one non-micro step on a stream (shared by the eager path and compiled plans)
static int enqueue_big(qtb_ctx *ctx, const StepGeom &g, int kind, const GettChoice &gc, const double2 *A, const double2 *B,
                       double2 *C, cudaStream_t s) {
    if (kind == KIND_GETT) {
        // compute-bound class on full 64 x 64 x 16 tiles: operand tiles by TMA when the leg placement allows it
        if (gc.cfg == 7 || gc.cfg >= 11) {
            GettTmaParams P;
            if (build_gett_tma(g, gc.swap, A, B, C, P)) {
                const unsigned nTiles = P.g.nTilesX * P.g.nTilesY;
                const unsigned grid = std::min<unsigned>(nTiles, (unsigned)ctx->numSMs);
                g_gettTmaFn<<<grid, TmaCfg::NT, TmaCfg::SMEM, s>>>(P);
                CU(cudaGetLastError());
                ctx->stats.launches++;
                ctx->stats.tma_launches++;
                return QTB_OK;
            }
        }
        GettParams p;
        build_gett(g, gc, A, B, C, p);
        return launch_gett(ctx, p, gc.cfg, s);
    }
    if (kind == KIND_REDUCE) return launch_reduce(ctx, g, A, B, C, s);
    if (kind == KIND_APPLY) return launch_apply(ctx, g, gc.swap, A, B, C, s);
    DevStep st;
    make_devstep(g, A, B, C, kind, st);
    return launch_generic(ctx, st, s);
}

// n independent all-micro plans in ONE launch (k_micro_batch: 512 threads): one CTA per plan, or -- when the plans are heavy and there are SMs
// to spare -- one thread-block cluster per plan (kernels.cuh: the CTAs of a cluster share every level's items).  The cluster size is
// the largest c <= 8 for which all n clusters are resident at once (cudaOccupancyMaxActiveClusters: one CTA per SM, a cluster has to
// fit one GPC).  Light plans stay on one CTA: a cluster barrier per level costs more than it buys (p=1 QAOA terms, 27 000 units each:
// 0.072 ms per evaluation on one CTA per term, 0.076 ms on clusters of three).  QTB_MICRO_CLUSTER=c forces a size, 1 turns clusters off.
static const long long MICRO_CLUSTER_MIN_UNITS = 100000;
static int micro_cluster_forced() {
    static int forced = -2;
    if (forced == -2) { const char *e = getenv("QTB_MICRO_CLUSTER"); forced = e ? std::max(1, std::min(8, atoi(e))) : -1; }
    return forced;
}
static int micro_cluster_size(qtb_ctx *ctx, int n, long long unitsPerPlan) {
    const int forced = micro_cluster_forced();
    if (forced > 0) return forced;
    if (unitsPerPlan < MICRO_CLUSTER_MIN_UNITS) return 1;
    static std::mutex mu;
    static std::unordered_map<long long, int> cache;                   // (device, n) -> size
    std::lock_guard<std::mutex> lk(mu);
    const long long key = ((long long)ctx->device << 32) | (unsigned)n;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int best = 1;
    for (int c = std::min(8, ctx->numSMs / std::max(1, n)); c >= 2; c--) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n * c)); cfg.blockDim = dim3(QTB_MICRO_THREADS_BATCH); cfg.dynamicSmemBytes = QTB_MICRO_SMEM_FOR(QTB_MICRO_THREADS_BATCH);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)c; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int active = 0;
        if (cudaOccupancyMaxActiveClusters(&active, k_micro_batch, &cfg) == cudaSuccess && active >= n) { best = c; break; }
        cudaGetLastError();
    }
    cache[key] = best;
    return best;
}
static int launch_micro_plans(qtb_ctx *ctx, int n, long long unitsTotal, const uint64_t *blobAddrDev, cudaStream_t s, const uint8_t *blobBase = nullptr) {
    const int c = micro_cluster_size(ctx, n, unitsTotal / std::max(1, n));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * c)); cfg.blockDim = dim3(QTB_MICRO_THREADS_BATCH); cfg.dynamicSmemBytes = QTB_MICRO_SMEM_FOR(QTB_MICRO_THREADS_BATCH); cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)c; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = c > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_micro_batch, blobBase, blobAddrDev);
    if (e != cudaSuccess && c > 1) {
        // a cluster that cannot be scheduled after all (the occupancy query is advisory): one CTA per plan always can
        cudaGetLastError();
        cfg.gridDim = dim3((unsigned)n); cfg.numAttrs = 0;
        e = cudaLaunchKernelEx(&cfg, k_micro_batch, blobBase, blobAddrDev);
    }
    CU(e);
    return QTB_OK;
}
// one grouped launch of a single plan segment / eager group: a long chain of tiny steps stays on one CTA of 1024 threads (GHZ-1000:
// 2 999 levels of one item, a cluster barrier per level would triple its time); a heavy group -- the 273 micro-steps of a config-2
// term are 1.2e6 units -- takes the batch kernel on a cluster
static int launch_micro_group(qtb_ctx *ctx, const uint8_t *blobBase, const uint64_t *offsetDev, long long units, cudaStream_t s) {
    if ((micro_cluster_forced() > 1 || units >= 4 * MICRO_CLUSTER_MIN_UNITS) && micro_cluster_size(ctx, 1, units) > 1)
        return launch_micro_plans(ctx, 1, units, offsetDev, s, blobBase);
    k_micro<<<1, QTB_MICRO_THREADS, QTB_MICRO_SMEM, s>>>(blobBase, offsetDev);
    CU(cudaGetLastError());
    return QTB_OK;
}

// ================================================================================================
// C ABI
extern "C" {

int qtb_abi_version(void) { return QTB_ABI_VERSION; }

const char *qtb_status_string(int s) {
    switch (s) {
        case QTB_OK: return "ok";
        case QTB_ERR_NO_DEVICE: return "no CUDA device (qtorch_b200 has no CPU fallback)";
        case QTB_ERR_INVALID: return "invalid argument";
        case QTB_ERR_EMPTY_INPUT: return "operand has no data";
        case QTB_ERR_OOM: return "out of device memory";
        case QTB_ERR_CUDA: return "CUDA error";
        case QTB_ERR_NCCL: return "NCCL error";
        case QTB_ERR_UNSUPPORTED: return "unsupported";
    }
    return "unknown status";
}
const char *qtb_last_error(void) { return g_lastError.c_str(); }

int qtb_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); if (count) *count = 0; return fail(QTB_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    if (count) *count = n;
    return QTB_OK;
}

static int ctx_init(qtb_ctx *ctx, int device) {
    ctx->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(QTB_ERR_NO_DEVICE, "device is not sm_100 (B200): kernels are built for sm_100a only");
    ctx->numSMs = prop.multiProcessorCount;
    if (const char *e = getenv("QTB_MICRO_LOG4")) ctx->microLog4 = std::max(0, std::min(10, atoi(e)));
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaMallocHost((void **)&ctx->ringHost, ctx->ringSize));
    CU(cudaMalloc((void **)&ctx->ringDev, ctx->ringSize));
    CU(cudaMallocHost((void **)&ctx->scalarPinned, 64));
    CU(cudaMalloc((void **)&ctx->zeroOffsetDev, 8));
    CU(cudaMalloc((void **)&ctx->reduceScratch, (size_t)REDUCE_MAX_BLOCKS * 16 * sizeof(double2)));
    CU(cudaMemset(ctx->zeroOffsetDev, 0, 8));
    CU(cudaEventCreateWithFlags(&ctx->ringEvent, cudaEventDisableTiming));
    CU(cudaFuncSetAttribute(k_micro, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QTB_MICRO_SMEM));
    CU(cudaFuncSetAttribute(k_micro_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QTB_MICRO_SMEM_FOR(QTB_MICRO_THREADS_BATCH)));
    for (auto &inst : g_gett) {
        CU(cudaFuncSetAttribute(inst.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inst.smem));
        int occ = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, inst.fn, inst.NT, inst.smem));
        inst.occ = std::max(1, occ);
    }
    CU(cudaFuncSetAttribute(g_gettTmaFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaCfg::SMEM));
    CU(cudaDeviceSynchronize());
    return QTB_OK;
}

int qtb_ctx_create(int device, qtb_ctx **out) {
    if (!out) return fail(QTB_ERR_INVALID, "null out");
    *out = nullptr;
    int n = 0;
    ST(qtb_device_count(&n));
    if (device < 0 || device >= n) return fail(QTB_ERR_INVALID, "device index out of range");
    qtb_ctx *ctx = new qtb_ctx_s();
    const int st = ctx_init(ctx, device);
    if (st != QTB_OK) {                      // give back whatever a half-built context holds (the error text survives)
        const std::string keep = g_lastError;
        qtb_ctx_destroy(ctx);
        g_lastError = keep;
        return st;
    }
    *out = ctx;
    return QTB_OK;
}

int qtb_ctx_destroy(qtb_ctx *ctx) {
    if (!ctx) return QTB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    if (ctx->commBuf) cudaFree(ctx->commBuf);
    for (auto &t : ctx->traceRecs) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    ctx->pool.destroy();
    for (cudaStream_t a : ctx->auxStreams) cudaStreamDestroy(a);
    for (cudaEvent_t e : ctx->auxEvents) cudaEventDestroy(e);
    if (ctx->forkEvent) cudaEventDestroy(ctx->forkEvent);
    for (qtb_scalar_read_s *r : ctx->freeReads) { cudaFreeHost(r->pinned); cudaEventDestroy(r->done); delete r; }
    if (ctx->ringHost) cudaFreeHost(ctx->ringHost);
    if (ctx->ringDev) cudaFree(ctx->ringDev);
    if (ctx->scalarPinned) cudaFreeHost(ctx->scalarPinned);
    if (ctx->zeroOffsetDev) cudaFree(ctx->zeroOffsetDev);
    if (ctx->reduceScratch) cudaFree(ctx->reduceScratch);
    if (ctx->batchOut) cudaFreeHost(ctx->batchOut);
    if (ctx->ringEvent) cudaEventDestroy(ctx->ringEvent);
    if (ctx->timer0) { cudaEventDestroy(ctx->timer0); cudaEventDestroy(ctx->timer1); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return QTB_OK;
}

int qtb_ctx_flush(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return flush_locked(ctx);
}
int qtb_ctx_sync(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(flush_locked(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    return QTB_OK;
}
void *qtb_ctx_stream(qtb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int qtb_tensor_alloc(qtb_ctx *ctx, int rank, qtb_tensor *out) {
    if (!ctx || !out) return fail(QTB_ERR_INVALID, "null argument");
    if (rank < 0 || rank > QTB_MAX_RANK) return fail(QTB_ERR_INVALID, "rank out of range");
    qtb_tensor t = new qtb_tensor_s();
    t->rank = rank;                      // device memory is bound lazily (first upload / first use as an output)
    *out = t;
    return QTB_OK;
}

int qtb_tensor_free(qtb_ctx *ctx, qtb_tensor t) {
    if (!t) return QTB_OK;
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (t->d && t->pooled) {
        // stream-ordered: deferred micro-steps may still reference the buffer, so it returns to the
        // pool right after the next flush is enqueued
        if (ctx->pending.empty() && ctx->pendingUploads.empty() && !ctx->held.active) ctx->pool.release(t->rank, t->d);
        else ctx->deferredFrees.push_back({t->rank, (void *)t->d});
    }
    delete t;
    return QTB_OK;
}

int qtb_tensor_rank(qtb_tensor t) { return t ? t->rank : -1; }
void *qtb_tensor_device_ptr(qtb_tensor t) { return t ? (void *)t->d : nullptr; }

int qtb_tensor_upload(qtb_ctx *ctx, qtb_tensor t, const double *host) {
    if (!ctx || !t || !host) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    ST(ensure_buffer(ctx, t));
    const size_t bytes = Pool::bytes(t->rank);
    if (t->rank <= 5 && !disable_micro()) {
        // rides in the next grouped launch as a level-0 copy item (one H2D for the whole batch)
        // Upload copies run at level 0 of the grouped launch: a buffer that a pending step still reads or writes must not be
        // overwritten there (re-uploading angles into a live handle), so such an upload starts a new batch.
        if (ctx->payload.size() + bytes > ((size_t)8 << 20) || ctx->pending.size() + ctx->pendingUploads.size() >= 8192 ||
            ctx->producedLevel.count(t->d) || ctx->readLevel.count(t->d)) ST(flush_locked(ctx));
        const size_t off = ctx->payload.size();
        ctx->payload.resize(off + bytes);
        memcpy(ctx->payload.data() + off, host, bytes);
        ctx->pendingUploads.push_back({t->d, off, (uint32_t)(bytes / 16)});
        ctx->producedLevel[t->d] = 0;
    } else {
        ST(flush_locked(ctx));
        CU(cudaMemcpyAsync(t->d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));       // pageable source: the caller may reuse it on return
    }
    ctx->stats.bytes_h2d += (long long)bytes;
    t->hasData = true;
    return QTB_OK;
}

int qtb_tensor_download(qtb_ctx *ctx, qtb_tensor t, double *host) {
    if (!ctx || !t || !host) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!t->hasData || !t->d) return fail(QTB_ERR_EMPTY_INPUT, "tensor has no data");
    ST(flush_locked(ctx));
    const size_t bytes = Pool::bytes(t->rank);
    CU(cudaMemcpyAsync(host, t->d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.bytes_d2h += (long long)bytes;
    return QTB_OK;
}

int qtb_read_scalar(qtb_ctx *ctx, qtb_tensor t, double out[2]) {
    if (!ctx || !t || !out) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!t->hasData || !t->d) return fail(QTB_ERR_EMPTY_INPUT, "tensor has no data");
    ST(flush_locked(ctx));
    CU(cudaMemcpyAsync(ctx->scalarPinned, t->d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    out[0] = ctx->scalarPinned[0]; out[1] = ctx->scalarPinned[1];
    ctx->stats.bytes_d2h += 16;
    return QTB_OK;
}

// queue a 16-byte device->host copy of *src on the ctx stream into a recycled pinned slot; the slot's event marks its arrival
static int scalar_read_enqueue(qtb_ctx *ctx, const double2 *src, qtb_scalar_read **out) {
    qtb_scalar_read *r = nullptr;
    if (!ctx->freeReads.empty()) { r = ctx->freeReads.back(); ctx->freeReads.pop_back(); }
    else {
        r = new qtb_scalar_read_s();
        if (cudaMallocHost((void **)&r->pinned, 16) != cudaSuccess || cudaEventCreateWithFlags(&r->done, cudaEventDisableTiming) != cudaSuccess) {
            if (r->pinned) cudaFreeHost(r->pinned);
            delete r;
            cudaGetLastError();
            return fail(QTB_ERR_OOM, "scalar read slot allocation failed");
        }
    }
    cudaError_t e = cudaMemcpyAsync(r->pinned, src, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(r->done, ctx->stream);
    if (e != cudaSuccess) { ctx->freeReads.push_back(r); return fail(QTB_ERR_CUDA, cudaGetErrorString(e)); }
    ctx->stats.bytes_d2h += 16;
    *out = r;
    return QTB_OK;
}

int qtb_read_scalar_begin(qtb_ctx *ctx, qtb_tensor t, qtb_scalar_read **out) {
    if (!ctx || !t || !out) return fail(QTB_ERR_INVALID, "null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!t->hasData || !t->d) return fail(QTB_ERR_EMPTY_INPUT, "tensor has no data");
    ST(flush_locked(ctx));
    return scalar_read_enqueue(ctx, t->d, out);
}

int qtb_read_scalar_end(qtb_ctx *ctx, qtb_scalar_read *r, double out[2]) {
    if (!ctx || !r) return fail(QTB_ERR_INVALID, "null argument");
    cudaError_t e = cudaEventSynchronize(r->done);           // outside the lock: other threads may keep enqueueing
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (out) { out[0] = r->pinned[0]; out[1] = r->pinned[1]; }
    ctx->freeReads.push_back(r);
    if (e != cudaSuccess) return fail(QTB_ERR_CUDA, cudaGetErrorString(e));
    return QTB_OK;
}

int qtb_contract(qtb_ctx *ctx, qtb_tensor a, qtb_tensor b, int k, const int *pos_a, const int *pos_b, qtb_tensor c) {
    if (!ctx || !a || !b || !c) return fail(QTB_ERR_INVALID, "null argument");
    if (k > 0 && (!pos_a || !pos_b)) return fail(QTB_ERR_INVALID, "null leg map");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!a->hasData || !b->hasData || !a->d || !b->d) return fail(QTB_ERR_EMPTY_INPUT, "operand has no data (Network.h:938-940)");
    StepGeom g;
    ST(make_geom(a->rank, b->rank, k, pos_a, pos_b, g));
    if (g.rC != c->rank) return fail(QTB_ERR_INVALID, "rank(C) != rank(A)+rank(B)-2k");
    if (c == a || c == b) return fail(QTB_ERR_INVALID, "output aliases an operand");
    if (a == b) return fail(QTB_ERR_INVALID, "a tensor cannot be contracted with itself (a step joins two distinct nodes, Network.h:715-716)");
    ST(ensure_device(ctx));
    ST(ensure_buffer(ctx, c));
    GettChoice gc{0, false};
    const int kind = choose_kind(g, gc, ctx->microLog4);
    ctx->stats.steps++;
    ctx->stats.units += (long long)g.units();
    if (kind == KIND_MICRO) {
        PendingStep ps;
        make_devstep(g, a->d, b->d, c->d, KIND_MICRO, ps.st);
        uint32_t lvl = 0;
        auto ia = ctx->producedLevel.find(a->d);
        if (ia != ctx->producedLevel.end()) lvl = std::max(lvl, ia->second + 1);
        auto ib = ctx->producedLevel.find(b->d);
        if (ib != ctx->producedLevel.end()) lvl = std::max(lvl, ib->second + 1);
        // an output buffer that pending steps still read (write-after-read) or write (write-after-write: out= re-used, or a
        // level-0 upload copy) is rewritten only after them
        auto ir = ctx->readLevel.find(c->d);
        if (ir != ctx->readLevel.end()) lvl = std::max(lvl, ir->second + 1);
        auto iw = ctx->producedLevel.find(c->d);
        if (iw != ctx->producedLevel.end()) lvl = std::max(lvl, iw->second + 1);
        ps.level = lvl;
        ctx->pending.push_back(ps);
        ctx->producedLevel[c->d] = lvl;
        for (const void *src : {(const void *)a->d, (const void *)b->d}) {
            auto it = ctx->readLevel.find(src);
            if (it == ctx->readLevel.end()) ctx->readLevel[src] = lvl; else it->second = std::max(it->second, lvl);
        }
        if (ctx->pending.size() >= 8192) ST(flush_locked(ctx));
    } else if (ctx->held.active && kind == KIND_REDUCE && (a->d == ctx->held.C || b->d == ctx->held.C) &&
               fusable_pair(ctx->held.g, KIND_GETT, ctx->held.gc, g, a->d == ctx->held.C)) {
        // the held DMMA step and this inner product run as ONE kernel; the intermediate is never materialised
        const bool tIsA = a->d == ctx->held.C;
        const qtb_ctx_s::Held h = ctx->held;
        ctx->held.active = false;
        ST(flush_locked(ctx));                                    // deferred micro-steps requested in between go first
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (ctx->trace) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, ctx->stream)); }
        ST(enqueue_fused(ctx, h.g, h.gc, h.A, h.B, g, tIsA, tIsA ? b->d : a->d, c->d, ctx->stream));
        (tIsA ? a : b)->hasData = false;          // the intermediate was never written: any later use of it reports EMPTY_INPUT
        if (ctx->trace) { CU(cudaEventRecord(e1, ctx->stream)); ctx->traceRecs.push_back({e0, e1, h.g.rA, h.g.rB, h.g.k, KIND_FUSED}); }
        for (auto &f : ctx->deferredFrees) ctx->pool.release(f.first, f.second);
        ctx->deferredFrees.clear();
    } else {
        ST(flush_locked(ctx));
        if (kind == KIND_GETT && fusion_enabled() && fused_variant_of(gc.cfg) >= 0 && g.rC >= 10) {
            // hold it back until the next request shows whether it can be fused (launched by the next flush at the latest)
            ctx->held.active = true; ctx->held.g = g; ctx->held.gc = gc; ctx->held.A = a->d; ctx->held.B = b->d; ctx->held.C = c->d;
        } else {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (ctx->trace) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, ctx->stream)); }
            ST(enqueue_big(ctx, g, kind, gc, a->d, b->d, c->d, ctx->stream));
            if (ctx->trace) { CU(cudaEventRecord(e1, ctx->stream)); ctx->traceRecs.push_back({e0, e1, g.rA, g.rB, g.k, kind}); }
        }
    }
    c->hasData = true;
    return QTB_OK;
}

// debugging aids (not part of the public header)
int qtb_debug_micro_layout(int rC, int k, int target, int *lg, int *lp) {          // tests/test_abi.py checks its invariants on the CPU
    uint32_t a = 0, b = 0;
    const uint32_t serial = micro_item_layout(1u << (2 * rC), 1u << (2 * k), (uint32_t)target, a, b);
    if (lg) *lg = (int)a;
    if (lp) *lp = (int)b;
    return (int)serial;
}
// prints the clock stamps the grouped launches of compiled plans left behind
int qtb_debug_dump_micro_timelines(int maxBlobs) {
    cudaDeviceSynchronize();
    int n = 0;
    for (const MicroTimeline &tl : g_microTimelines) {
        if (n++ >= maxBlobs) break;
        std::vector<unsigned long long> h(tl.nLevels + 2);
        if (cudaMemcpy(h.data(), tl.dev, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); continue; }
        fprintf(stderr, "micro blob %d: %u levels, %llu cycles\n", n - 1, tl.nLevels, h[tl.nLevels] - h[0]);
        for (uint32_t l = 0; l < tl.nLevels; l++)
            fprintf(stderr, "  level %3u: %8llu cycles  %4u items  max serial %5u MACs\n", l, h[l + 1] - h[l], tl.items[l], tl.maxSerial[l]);
    }
    return n;
}

// ---- stats / trace ------------------------------------------------------------------------------
int qtb_ctx_stats(qtb_ctx *ctx, qtb_stats *out) {
    if (!ctx || !out) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    *out = ctx->stats;
    out->pool_bytes_reserved = ctx->pool.reserved;
    out->pool_bytes_peak_live = ctx->pool.peakLive;
    return QTB_OK;
}
int qtb_ctx_reset_stats(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->stats = qtb_stats{};
    return QTB_OK;
}
int qtb_ctx_timer_start(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    if (!ctx->timer0) { CU(cudaEventCreate(&ctx->timer0)); CU(cudaEventCreate(&ctx->timer1)); }
    CU(cudaEventRecord(ctx->timer0, ctx->stream));
    return QTB_OK;
}
int qtb_ctx_timer_stop(qtb_ctx *ctx, float *ms) {
    if (!ctx || !ms) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->timer0) return fail(QTB_ERR_INVALID, "timer not started");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    CU(cudaEventRecord(ctx->timer1, ctx->stream));
    CU(cudaEventSynchronize(ctx->timer1));
    CU(cudaEventElapsedTime(ms, ctx->timer0, ctx->timer1));
    return QTB_OK;
}
int qtb_ctx_set_micro_limit(qtb_ctx *ctx, int log4Units) {
    if (!ctx || log4Units < 0 || log4Units > 10) return fail(QTB_ERR_INVALID, "micro limit must be 0..10 (4^n complex MACs)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(flush_locked(ctx));
    ctx->microLog4 = log4Units;
    return QTB_OK;
}
int qtb_ctx_get_micro_limit(qtb_ctx *ctx) { return ctx ? ctx->microLog4 : -1; }
int qtb_ctx_trace_enable(qtb_ctx *ctx, int on) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->trace = on != 0;
    return QTB_OK;
}
int qtb_ctx_trace_read(qtb_ctx *ctx, qtb_step_trace *out, int maxEntries, int *n) {
    if (!ctx || !n) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(flush_locked(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    int cnt = 0;
    for (auto &t : ctx->traceRecs) {
        if (out && cnt < maxEntries) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, t.e0, t.e1));
            out[cnt].rank_a = t.rA; out[cnt].rank_b = t.rB; out[cnt].k = t.k; out[cnt].kernel = t.kernel; out[cnt].ms = ms;
        }
        cnt++;
        cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
    }
    ctx->traceRecs.clear();
    *n = std::min(cnt, maxEntries);
    return QTB_OK;
}

// ---- NCCL ------------------------------------------------------------------------------------------
int qtb_comm_unique_id(char id[QTB_UNIQUE_ID_BYTES]) {
    ST(load_nccl());
    NcclId nid; memset(&nid, 0, sizeof(nid));
    int r = g_nccl.GetUniqueId(&nid);
    if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    memcpy(id, nid.b, QTB_UNIQUE_ID_BYTES);
    return QTB_OK;
}
int qtb_comm_init(qtb_ctx *ctx, int nRanks, int rank, const char id[QTB_UNIQUE_ID_BYTES]) {
    if (!ctx || !id) return fail(QTB_ERR_INVALID, "null argument");
    ST(load_nccl());
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    NcclId nid; memcpy(nid.b, id, QTB_UNIQUE_ID_BYTES);
    int r = g_nccl.CommInitRank(&ctx->comm, nRanks, nid, rank);
    if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    ctx->nRanks = nRanks; ctx->rank = rank;
    return QTB_OK;
}
int qtb_comm_destroy(qtb_ctx *ctx) {
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
    return QTB_OK;
}
int qtb_allreduce_sum(qtb_ctx *ctx, double *host, int nComplex) {
    if (!ctx || !host || nComplex < 0) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->comm) return fail(QTB_ERR_NCCL, "communicator not initialised (qtb_comm_init)");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    const size_t n = (size_t)nComplex * 2;
    if (n > ctx->commBufElems) {
        if (ctx->commBuf) CU(cudaFree(ctx->commBuf));
        CU(cudaMalloc((void **)&ctx->commBuf, std::max<size_t>(n, 1024) * 8));
        ctx->commBufElems = std::max<size_t>(n, 1024);
    }
    CU(cudaMemcpyAsync(ctx->commBuf, host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int r = g_nccl.AllReduce(ctx->commBuf, ctx->commBuf, n, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
    if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    CU(cudaMemcpyAsync(host, ctx->commBuf, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return QTB_OK;
}

// the same reduction on a buffer that already lives on the device: one in-stream ncclAllReduce, no host staging, no wait
int qtb_allreduce_sum_device(qtb_ctx *ctx, void *dev, int nComplex) {
    if (!ctx || !dev || nComplex < 0) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->comm) return fail(QTB_ERR_NCCL, "communicator not initialised (qtb_comm_init)");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    int r = g_nccl.AllReduce(dev, dev, (size_t)nComplex * 2, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
    if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return QTB_OK;
}

}  // extern "C"

#include "plan.inl"
#include "sliced.inl"
#include "batch.inl"
