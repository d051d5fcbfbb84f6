/* oracle/contract_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C11) of the reference qTorch pairwise contraction:
 *   Network::ContractIndices   /root/reference/src/Network.h:876-971   (arithmetic, loop order)
 *   Network::ContractNodes     /root/reference/src/Network.h:715-864   (leg bookkeeping)
 *   Node::Index / Node::Access /root/reference/src/Node.h:168-194      (little-endian base-4 layout)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (qtorch_b200/) never links it and
 * has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit against the
 * unmodified reference (oracle/_ref/ref_harness, mode "step") and against the golden fixtures
 * in tests/golden/ that were generated from the reference by tests/golden/make_golden.py.
 *
 * Conventions restated from the reference:
 *   - every leg has dimension 4; element i of a rank-r tensor has digit d_j = (i >> 2j) & 3 on
 *     leg j, i.e. leg 0 is the fastest (Node.h:178-186);
 *   - shared legs are listed in A-leg order: pair j = (posA[j], posB[j]) with posA increasing
 *     (Network.h:739-758);
 *   - C's legs are A's free legs in A order followed by B's free legs in B order
 *     (Network.h:743-747, 762-768, 809-812);
 *   - C[c] = sum_{s=0}^{4^k-1} A[..] * B[..] where digit i of s drives shared pair (k-1-i),
 *     i.e. the LAST shared pair varies fastest (Network.h:912-916); accumulation in that order
 *     starting from 0 (Node.h:112-113 zero-fills C), plain complex multiply then add, no FMA
 *     (compile with -ffp-contract=off to keep it bit-identical to the reference on x86-64);
 *   - the "float op" counter adds 4^(rC+k) per step (Network.h:884-885).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#pragma STDC FP_CONTRACT OFF

/* worker threads for qto_contract (contiguous chunks of the output range, like the
 * reference's std::thread fan-out, Network.h:941-960); per-element arithmetic order is
 * unchanged, so results do not depend on the thread count. */
static int g_threads = 1;
void qto_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int qto_get_threads(void) { return g_threads; }

typedef struct { double re, im; } qto_c64;

/* Derive the free-leg lists the way ContractNodes does.  Returns rC, or -1 on bad input. */
int qto_free_legs(int rA, int rB, int k, const int *posA, const int *posB,
                  int *freeA, int *nFreeA, int *freeB, int *nFreeB)
{
    int usedA[32] = {0}, usedB[32] = {0};
    if (rA < 0 || rB < 0 || rA > 31 || rB > 31 || k < 0 || k > rA || k > rB) return -1;
    for (int j = 0; j < k; j++) {
        if (posA[j] < 0 || posA[j] >= rA || posB[j] < 0 || posB[j] >= rB) return -1;
        if (usedA[posA[j]] || usedB[posB[j]]) return -1;
        if (j > 0 && posA[j] <= posA[j - 1]) return -1;   /* A-leg order, Network.h:739-749 */
        usedA[posA[j]] = 1; usedB[posB[j]] = 1;
    }
    int na = 0, nb = 0;
    for (int i = 0; i < rA; i++) if (!usedA[i]) freeA[na++] = i;
    for (int i = 0; i < rB; i++) if (!usedB[i]) freeB[nb++] = i;
    *nFreeA = na; *nFreeB = nb;
    return na + nb;
}

/* units (the reference's getNumFloatOps increment) for one step */
long long qto_step_units(int rA, int rB, int k) { return 1LL << (2 * (rA + rB - 2 * k + k)); }

typedef struct {
    const qto_c64 *A, *B; qto_c64 *C;
    const uint64_t *sA, *sB; uint64_t K, c0, c1;
    const int *freeA, *freeB; int na, nb;
} qto_job;

static void *qto_worker(void *arg)
{
    const qto_job *j = (const qto_job *)arg;
    const qto_c64 *A = j->A, *B = j->B; qto_c64 *C = j->C;
    const uint64_t *sA = j->sA, *sB = j->sB; const uint64_t K = j->K, c0 = j->c0, c1 = j->c1;
    const int *freeA = j->freeA, *freeB = j->freeB; const int na = j->na, nb = j->nb;
    for (uint64_t c = c0; c < c1; c++) {
        /* output digit i <-> toNotSumOn[i] (Network.h:902-908) */
        uint64_t baseA = 0, baseB = 0;
        for (int i = 0; i < na; i++) baseA += ((c >> (2 * i)) & 3) << (2 * freeA[i]);
        for (int i = 0; i < nb; i++) baseB += ((c >> (2 * (na + i))) & 3) << (2 * freeB[i]);
        double accr = 0.0, acci = 0.0;
        for (uint64_t s = 0; s < K; s++) {
            const qto_c64 a = A[baseA + sA[s]], b = B[baseB + sB[s]];
            /* std::complex<double> operator*: (ac - bd, ad + bc), then += (Network.h:931) */
            const double pr = a.re * b.re - a.im * b.im;
            const double pi = a.re * b.im + a.im * b.re;
            accr += pr; acci += pi;
        }
        C[c].re = accr; C[c].im = acci;
    }
    return NULL;
}

/* The contraction.  A has 4^rA elements, B 4^rB, C 4^(rA+rB-2k); C is overwritten.
 * Returns 0 on success, -1 on bad leg maps. */
int qto_contract(const qto_c64 *A, int rA, const qto_c64 *B, int rB, int k,
                 const int *posA, const int *posB, qto_c64 *C)
{
    int freeA[32], freeB[32], na, nb;
    int rC = qto_free_legs(rA, rB, k, posA, posB, freeA, &na, freeB, &nb);
    if (rC < 0) return -1;
    const uint64_t K = 1ULL << (2 * k), NC = 1ULL << (2 * rC);

    /* offsets contributed by the summed counter s (Network.h:912-916: digit i <-> pair k-1-i) */
    uint64_t *sA = (uint64_t *)malloc(sizeof(uint64_t) * K);
    uint64_t *sB = (uint64_t *)malloc(sizeof(uint64_t) * K);
    for (uint64_t s = 0; s < K; s++) {
        uint64_t oa = 0, ob = 0;
        for (int i = 0; i < k; i++) {
            uint64_t d = (s >> (2 * i)) & 3;
            oa += d << (2 * posA[k - 1 - i]);
            ob += d << (2 * posB[k - 1 - i]);
        }
        sA[s] = oa; sB[s] = ob;
    }
    int nt = (NC * K >= (1u << 16)) ? g_threads : 1;
    if ((uint64_t)nt > NC) nt = (int)NC;
    qto_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < nt; t++) {
        qto_job jb = {A, B, C, sA, sB, K, NC * t / nt, NC * (t + 1) / nt, freeA, freeB, na, nb};
        jobs[t] = jb;
        if (t + 1 < nt) pthread_create(&th[t], NULL, qto_worker, &jobs[t]);
    }
    qto_worker(&jobs[nt - 1]);
    for (int t = 0; t + 1 < nt; t++) pthread_join(th[t], NULL);
    free(sA); free(sB);
    return 0;
}

/* The final-value rule of Network.h:961-969, restated for host-side tests:
 * returns 1 if mFinalVal must be replaced by C[0] after a rank-0 result. */
int qto_final_value_rule(double cur_re, double cur_im, int rA, int rB)
{
    return (fabs(cur_re) <= 1.0e-30 && fabs(cur_im) <= 1.0e-30) || (rA == 0 && rB == 0);
}
