// qtorch_b200/csrc/step.h -- the device-side description of one pairwise contraction step.
//
// A step is exactly the information Network::ContractNodes hands to Network::ContractIndices
// (/root/reference/src/Network.h:837: toNotSumOn, toSumOn) re-expressed as bit shifts into the
// little-endian base-4 element index of each operand (Node.h:178-186).
#pragma once
#include <stdint.h>

#define QTB_MAXR 16

// kernel families (DevStep::kind, qtb_step_trace::kernel)
enum StepKind { KIND_MICRO = 0, KIND_THREAD = 1, KIND_GETT = 2, KIND_WARP = 3, KIND_COPY = 4, KIND_REDUCE = 5, KIND_FUSED = 6, KIND_APPLY = 7 };

struct DevStep {
    const double2 *A;
    const double2 *B;
    double2 *C;
    uint8_t rA, rB, k, rC;
    uint8_t nfa, nfb;                 // free legs of A / of B; rC = nfa + nfb
    uint8_t kind;                     // kernel family chosen by the host (see engine.cu)
    uint8_t pad0;
    // output digit i (0 <= i < rC): bit shift of that digit inside A (i < nfa) or B (i >= nfa)
    uint8_t shFree[QTB_MAXR * 2];
    // summed digit i (0 <= i < k): bit shifts inside A and inside B.  Digit 0 is the LAST shared
    // pair, like the reference's inner counter (Network.h:912-916), so sequential accumulation
    // visits the terms in the reference's order.
    uint8_t shSumA[QTB_MAXR];
    uint8_t shSumB[QTB_MAXR];
};

// Host-side geometry of a step (leg lists in the reference's conventions).
struct StepGeom {
    int rA, rB, k, rC, nfa, nfb;
    int posA[QTB_MAXR], posB[QTB_MAXR];      // shared pairs, posA increasing
    int freeA[QTB_MAXR], freeB[QTB_MAXR];    // free leg positions, increasing
    unsigned long long units() const { return 1ULL << (2 * (rC + k)); }
};
