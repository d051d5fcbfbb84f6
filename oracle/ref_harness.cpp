// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference qTorch headers (compiled where they lie under
// /root/reference/src, via -I) through their own public API and prints results at full
// precision.  It is the "real reference" arm of the oracle: it pins the C restatement in
// oracle/contract_oracle.c, generates the golden fixtures under tests/golden/, and is the
// CPU baseline that bench.py times (`cpu_baseline.kind == "reference"`).
//
// The reference is header-only with non-inline definitions and globals in headers
// (src/Network.h:50-51, src/Timer.h:40-60), so exactly ONE translation unit may include it:
// this file.  Built by oracle/Makefile into oracle/_ref/ref_harness (git-ignored).
//
// Every result line is prefixed "@@" so callers can ignore the reference's own chatter.
//
// Modes
//   step  rA rB k pA.. pB.. A.bin B.bin C.bin      one pairwise contraction through
//                                                   Network::ContractNodes (src/Network.h:715)
//   stepbench rA rB k pA.. pB.. threads            the same single step on seeded pseudo-random tensors built in memory, timed
//                                                   (the CPU arm's sample of the "large tensor" step class; nothing written)
//   gate  NAME [angle|file]                         dump a gate/measurement tensor table (src/Node.h:197-898)
//   lg    qasm measure qbb.out reduce threads       ReduceCircuit + LineGraph::LGContract on a frozen ordering
//   qbb   qasm measure seconds cnf out stats reduce run quickbb_64 (must be on PATH) -> ordering files
//   seq   qasm measure plan.txt threads [budget]    replay "(a b)" pairs via ContractNodes, per-step timing;
//                                                   stops once `budget` units have been executed (bounded sample)
//   stoch qasm measure threads                      ContractionTools::Contract(Stochastic), prints the plan
//   cost  qasm measure pValue seed threads          ContractionTools::Contract(CostContractSimple, pValue) with the
//                                                   generator seeded (see the access note below), prints the plan
//   inp   script.inp                                leviParser on a ">type key value" script, dumps its four maps
//   user  qasm measure seqfile                      ContractUserDefinedSequenceOfWires
//   maxcut graph.dgf p outdir b1..bp g1..gp         the body of F_p (src/maxcut.cpp:162-204) for fixed angles: per-edge
//                                                   light-cone circuits written with the reference's own emitters
//                                                   (outdir/term<i>.qasm), contracted stochastically, summed
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <complex>
#include <chrono>

// ContractionTools seeds its private std::mt19937 from std::random_device (src/ContractionTools.h:61,97), so its random
// planners cannot be replayed.  To pin the host mirror's restatement of CostContractSimple draw by draw, the `cost` mode
// assigns that generator a fixed seed; the access override below is the only liberty taken with the reference, it is
// compile-time, confined to this test harness, and does not touch the sources (every standard header the reference
// pulls in is included BEFORE the override so that only the reference's own classes are affected).
#include <algorithm>
#include <array>
#include <csignal>
#include <exception>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <regex>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <sys/stat.h>
#define private public
#include "ContractionTools.h"   // reference header (pulls Network.h, Node.h, LineGraph.h ...)
#undef private
#include "leviParser.hpp"       // reference .inp reader (the `inp` mode)
#include "maxcut.h"             // reference QAOA helpers (ExtraData, circuit emitters); <nlopt.hpp> = oracle/stubs

using namespace qtorch;
typedef std::complex<double> cplx;

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static std::vector<cplx> read_bin(const char *path, size_t n) {
    std::vector<cplx> v(n);
    FILE *f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(cplx), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

static void print_plan(const std::shared_ptr<Network> &net) {
    // plan fingerprint = mCreatedFrom pairs in mAllNodes order (idiom of src/preprocess.h:40-45)
    printf("@@plan");
    for (const auto &n : net->GetAllNodes()) {
        if (!(n->mCreatedFrom.first == 0 && n->mCreatedFrom.second == 0))
            printf(" %d,%d", n->mCreatedFrom.first, n->mCreatedFrom.second);
    }
    printf("\n@@nodes %zu\n", net->GetAllNodes().size());
}

static void print_value(const char *tag, cplx v) { printf("@@%s %.17g %.17g\n", tag, v.real(), v.imag()); }

static int mode_step(int argc, char **argv) {
    int a = 2;
    int rA = atoi(argv[a++]), rB = atoi(argv[a++]), k = atoi(argv[a++]);
    std::vector<int> pA(k), pB(k);
    for (int i = 0; i < k; i++) pA[i] = atoi(argv[a++]);
    for (int i = 0; i < k; i++) pB[i] = atoi(argv[a++]);
    const char *fa = argv[a++], *fb = argv[a++], *fc = argv[a++];
    auto A = std::make_shared<Node>(rA);
    auto B = std::make_shared<Node>(rB);
    A->GetTensorVals() = read_bin(fa, (size_t)1 << (2 * rA));
    B->GetTensorVals() = read_bin(fb, (size_t)1 << (2 * rB));
    std::vector<std::shared_ptr<Wire>> wa(rA), wb(rB);
    for (int i = 0; i < k; i++) {
        auto w = std::make_shared<Wire>(A, B, 0);
        wa[pA[i]] = w; wb[pB[i]] = w;
    }
    for (int i = 0; i < rA; i++) if (!wa[i]) wa[i] = std::make_shared<Wire>(A, nullptr, 0);
    for (int i = 0; i < rB; i++) if (!wb[i]) wb[i] = std::make_shared<Wire>(nullptr, B, 0);
    for (auto &w : wa) A->GetWires().push_back(w);
    for (auto &w : wb) B->GetWires().push_back(w);
    Network net;
    net.SetNumThreads(8);
    std::shared_ptr<Node> C = net.ContractNodes(A, B, 1000);
    if (!C) { printf("@@rejected\n"); return 0; }
    FILE *f = fopen(fc, "wb");
    fwrite(C->GetTensorVals().data(), sizeof(cplx), C->GetTensorVals().size(), f);
    fclose(f);
    printf("@@rank %d\n@@flops %lld\n", C->mRank, net.getNumFloatOps());
    return 0;
}

// one step of a given shape on in-memory pseudo-random operands through Network::ContractNodes, timed
static int mode_stepbench(int argc, char **argv) {
    int a = 2;
    int rA = atoi(argv[a++]), rB = atoi(argv[a++]), k = atoi(argv[a++]);
    std::vector<int> pA(k), pB(k);
    for (int i = 0; i < k; i++) pA[i] = atoi(argv[a++]);
    for (int i = 0; i < k; i++) pB[i] = atoi(argv[a++]);
    const int threads = atoi(argv[a++]);
    auto A = std::make_shared<Node>(rA);
    auto B = std::make_shared<Node>(rB);
    unsigned long long x = 88172645463325252ull;          // xorshift64: cheap, deterministic filler
    auto fill = [&x](std::vector<cplx> &v) {
        for (auto &e : v) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            const double re = (double)(x & 0xfffff) / 1048576.0 - 0.5;
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            e = cplx(re, (double)(x & 0xfffff) / 1048576.0 - 0.5);
        }
    };
    fill(A->GetTensorVals());
    fill(B->GetTensorVals());
    std::vector<std::shared_ptr<Wire>> wa(rA), wb(rB);
    for (int i = 0; i < k; i++) {
        auto w = std::make_shared<Wire>(A, B, 0);
        wa[pA[i]] = w; wb[pB[i]] = w;
    }
    for (int i = 0; i < rA; i++) if (!wa[i]) wa[i] = std::make_shared<Wire>(A, nullptr, 0);
    for (int i = 0; i < rB; i++) if (!wb[i]) wb[i] = std::make_shared<Wire>(nullptr, B, 0);
    for (auto &w : wa) A->GetWires().push_back(w);
    for (auto &w : wb) B->GetWires().push_back(w);
    Network net;
    net.SetNumThreads(threads);
    const double t0 = now_s();
    std::shared_ptr<Node> C = net.ContractNodes(A, B, 1000);
    const double dt = now_s() - t0;
    if (!C) { printf("@@rejected\n"); return 0; }
    printf("@@rank %d\n@@flops %lld\n@@seconds %.6f\n", C->mRank, net.getNumFloatOps(), dt);
    print_value("probe", C->GetTensorVals()[C->GetTensorVals().size() / 3]);
    return 0;
}

static int mode_gate(int argc, char **argv) {
    std::string name = argv[2];
    std::shared_ptr<Node> n;
    double ang = argc > 3 ? atof(argv[3]) : 0.0;
    if (name == "CNOT") n = std::make_shared<CNOTNode>();
    else if (name == "SWAP") n = std::make_shared<SwapNode>();
    else if (name == "H") n = std::make_shared<HNode>();
    else if (name == "X") n = std::make_shared<XNode>();
    else if (name == "Y") n = std::make_shared<YNode>();
    else if (name == "Z") n = std::make_shared<ZNode>();
    else if (name == "Rx") n = std::make_shared<RxNode>(ang);
    else if (name == "Ry") n = std::make_shared<RyNode>(ang);
    else if (name == "Rz") n = std::make_shared<RzNode>(ang);
    else if (name == "PHASE") n = std::make_shared<PhaseNode>(ang);
    else if (name == "CZ") n = std::make_shared<CZNode>();
    else if (name == "CRk") n = std::make_shared<CRkNode>((int)ang);
    else if (name == "CPHASE") n = std::make_shared<CPhaseNode>(ang);
    else if (name == "ZeroState") n = std::make_shared<ZeroStateNode>();
    else if (name == "Trace") n = std::make_shared<TraceNode>();
    else if (name == "XMeasure") n = std::make_shared<XMeasure>();
    else if (name == "YMeasure") n = std::make_shared<YMeasure>();
    else if (name == "ZMeasure") n = std::make_shared<ZMeasure>();
    else if (name == "ProjectOne") n = std::make_shared<ProjectOne>();
    else if (name == "ProjectZero") n = std::make_shared<ProjectZero>();
    else if (name == "def1") n = std::make_shared<ArbitraryOneQubitNode>(argv[3], "u1");
    else if (name == "def2") n = std::make_shared<ArbitraryTwoQubitNode>(argv[3], "u2");
    else { fprintf(stderr, "unknown gate %s\n", name.c_str()); return 2; }
    printf("@@gate %s %d", name.c_str(), n->mRank);
    for (const auto &v : n->GetTensorVals()) printf(" %.17g %.17g", v.real(), v.imag());
    printf("\n");
    return 0;
}

static int mode_lg(int argc, char **argv) {
    const char *qasm = argv[2], *meas = argv[3], *qbb = argv[4];
    int reduce = atoi(argv[5]), threads = atoi(argv[6]);
    auto net = std::make_shared<Network>(qasm, meas);
    net->SetNumThreads(threads);
    double t0 = now_s();
    if (reduce) net->ReduceCircuit();
    LineGraph lg(net);
    lg.SetQBBOutFiles("/dev/null", qbb, "/dev/null");
    bool ok = false;
    try { ok = lg.LGContract(); } catch (std::exception &e) { printf("@@exception %s\n", e.what()); }
    double t1 = now_s();
    printf("@@ok %d\n", (int)ok);
    print_value("value", net->GetFinalValue());
    printf("@@flops %lld\n@@seconds %.6f\n", net->getNumFloatOps(), t1 - t0);
    print_plan(net);
    return 0;
}

static int mode_qbb(int argc, char **argv) {
    const char *qasm = argv[2], *meas = argv[3];
    int secs = atoi(argv[4]);
    const char *cnf = argv[5], *out = argv[6], *stats = argv[7];
    int reduce = atoi(argv[8]);
    auto net = std::make_shared<Network>(qasm, meas);
    if (reduce) net->ReduceCircuit();
    LineGraph lg(net);
    lg.SetQBBOutFiles(cnf, out, stats);
    lg.runQuickBB(secs, nullptr, true);
    printf("@@qbb done\n");
    return 0;
}

static int mode_seq(int argc, char **argv) {
    const char *qasm = argv[2], *meas = argv[3], *planf = argv[4];
    int threads = atoi(argv[5]);
    double budget = argc > 6 ? atof(argv[6]) : 0.0;   // units; 0 = whole plan
    std::vector<std::pair<int, int>> plan;
    FILE *f = fopen(planf, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", planf); return 2; }
    int a, b;
    while (fscanf(f, "%d %d", &a, &b) == 2) plan.push_back({a, b});
    fclose(f);
    auto net = std::make_shared<Network>(qasm, meas);
    net->SetNumThreads(threads);
    double total = 0.0;
    long long done_units = 0;
    size_t steps = 0;
    for (const auto &p : plan) {
        if (budget > 0 && (double)done_units >= budget) break;
        long long before = net->getNumFloatOps();
        auto A = net->GetAllNodes()[p.first];
        auto B = net->GetAllNodes()[p.second];
        int ra = A->mRank, rb = B->mRank;
        double t0 = now_s();
        auto C = net->ContractNodes(A, B, 100);
        double dt = now_s() - t0;
        long long u = net->getNumFloatOps() - before;
        total += dt; done_units += u; steps++;
        printf("@@step %zu %d %d %d %lld %.9f\n", steps - 1, ra, rb, C ? C->mRank : 0, u, dt);
    }
    printf("@@done %d\n", (int)net->IsDone());
    print_value("value", net->GetFinalValue());
    printf("@@flops %lld\n@@steps %zu\n@@seconds %.6f\n", done_units, steps, total);
    return 0;
}

static int mode_stoch(int argc, char **argv) {
    const char *qasm = argv[2], *meas = argv[3];
    int threads = atoi(argv[4]);
    ContractionTools p(qasm, meas, threads);
    auto net = p.Contract(Stochastic);
    print_value("value", p.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    return 0;
}

static int mode_cost(int argc, char **argv) {
    const int pValue = atoi(argv[4]);
    const unsigned seed = (unsigned)strtoul(argv[5], nullptr, 10);
    ContractionTools p(argv[2], argv[3], argc > 6 ? atoi(argv[6]) : 8);
    p.mRandGen = std::mt19937(seed);
    auto net = p.Contract(CostContractSimple, pValue);
    print_value("value", p.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    return 0;
}

// inp script.inp: the reference's leviParser on a script, one line per map entry (same format as qtb_harness inp)
static int mode_inp(int argc, char **argv) {
    leviParser script;
    const bool opened = script.readInputFile(argv[2]);
    printf("@@opened %d\n", opened ? 1 : 0);
    for (const auto &kv : script.mapString) printf("@@string %s %s\n", kv.first.c_str(), kv.second.c_str());
    for (const auto &kv : script.mapBool) printf("@@bool %s %d\n", kv.first.c_str(), kv.second ? 1 : 0);
    for (const auto &kv : script.mapInt) printf("@@int %s %d\n", kv.first.c_str(), kv.second);
    for (const auto &kv : script.mapDouble) printf("@@double %s %.17g\n", kv.first.c_str(), kv.second);
    return 0;
}

static int mode_user(int argc, char **argv) {
    ContractionTools p(argv[2], argv[3]);
    auto net = p.ContractUserDefinedSequenceOfWires(argv[4]);
    print_value("value", p.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    return 0;
}

static int mode_maxcut(int argc, char **argv) {
    const char *graph = argv[2];
    const int p = atoi(argv[3]);
    const std::string outdir = argv[4];
    std::vector<double> bg(2 * p);
    for (int i = 0; i < 2 * p; i++) bg[i] = atof(argv[5 + i]);
    ExtraData e(p, graph);
    double fp = 0.0;
    printf("@@edges %zu\n", e.pairs.size());
    for (size_t c = 0; c < e.pairs.size(); c++) {
        const std::string qasm = outdir + "/term" + std::to_string(c) + ".qasm", meas = outdir + "/term" + std::to_string(c) + ".meas";
        {
            std::ofstream q(qasm), m(meas);
            const int nq = e.qubitsNeeded[c];
            q << nq << std::endl;
            outputInitialPlusStateToFile(q, nq);
            applyU_CsThenU_Bs(e.realIterations[c], e.p, bg, nq, q);
            for (int i = 0; i < nq; i++) m << ((i == 0 || i == 1) ? "Z " : "T ");
        }
        ContractionTools qc(qasm, meas);
        auto net = qc.Contract(Stochastic);
        printf("@@term %zu %d %d %d %.17g %.17g %lld\n", c, e.pairs[c].first, e.pairs[c].second, e.qubitsNeeded[c], qc.GetFinalVal().real(),
               qc.GetFinalVal().imag(), net->getNumFloatOps());
        fp += 0.5 * (1.0 - qc.GetFinalVal().real());
    }
    printf("@@fp %.17g\n", fp);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_harness <mode> ...\n"); return 2; }
    std::string m = argv[1];
    try {
        if (m == "step") return mode_step(argc, argv);
        if (m == "stepbench") return mode_stepbench(argc, argv);
        if (m == "gate") return mode_gate(argc, argv);
        if (m == "lg") return mode_lg(argc, argv);
        if (m == "qbb") return mode_qbb(argc, argv);
        if (m == "seq") return mode_seq(argc, argv);
        if (m == "stoch") return mode_stoch(argc, argv);
        if (m == "cost") return mode_cost(argc, argv);
        if (m == "inp") return mode_inp(argc, argv);
        if (m == "user") return mode_user(argc, argv);
        if (m == "maxcut") return mode_maxcut(argc, argv);
    } catch (std::exception &e) {
        printf("@@exception %s\n", e.what());
        return 1;
    }
    fprintf(stderr, "unknown mode %s\n", m.c_str());
    return 2;
}
