// qtorch_b200/apps/qtb_harness.cpp -- full-precision driver of the B200 host mirror, line-compatible with
// oracle/ref_harness.cpp (same modes, same "@@" output lines) so the parity tests can diff the two.
// Extra mode: `plan` prints every executed step with its leg maps (the input of qtb_plan_create).
// With QTORCH_PLAN_ONLY=1 no device is touched: only the plan / bookkeeping lines are meaningful.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/qtorch.hpp"

typedef std::complex<double> cplx;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void print_value(const char *tag, cplx v) { printf("@@%s %.17g %.17g\n", tag, v.real(), v.imag()); }

static void print_plan(const std::shared_ptr<Network> &net) {
    printf("@@plan");
    for (const auto &n : net->GetAllNodes())
        if (!(n->mCreatedFrom.first == 0 && n->mCreatedFrom.second == 0)) printf(" %d,%d", n->mCreatedFrom.first, n->mCreatedFrom.second);
    printf("\n@@nodes %zu\n", net->GetAllNodes().size());
}

// "@@pstep a b c rankA rankB rankC k posA.. posB.."
static void print_steps(const std::shared_ptr<Network> &net) {
    printf("@@inputs %d", net->GetNumOriginalNodes());
    for (int i = 0; i < net->GetNumOriginalNodes(); i++) printf(" %d", net->GetAllNodes()[i]->mRank);
    printf("\n");
    for (const auto &r : net->GetPlan()) {
        printf("@@pstep %d %d %d %d %d %d %zu", r.a, r.b, r.c, r.rankA, r.rankB, r.rankC, r.posA.size());
        for (int p : r.posA) printf(" %d", p);
        for (int p : r.posB) printf(" %d", p);
        printf("\n");
    }
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

static int mode_lg(int argc, char **argv, bool steps) {
    const char *qasm = argv[2], *meas = argv[3], *qbb = argv[4];
    int reduce = atoi(argv[5]);
    auto net = std::make_shared<Network>(qasm, meas);
    double t0 = now_s();
    if (reduce) net->ReduceCircuit();
    LineGraph lg(net);
    lg.SetQBBOutFiles("/dev/null", qbb, "/dev/null");
    bool ok = false;
    try { ok = lg.LGContract(); } catch (std::exception &e) { printf("@@exception %s\n", e.what()); }
    cplx v = net->GetFinalValue();
    double t1 = now_s();
    printf("@@ok %d\n", (int)ok);
    print_value("value", v);
    printf("@@flops %lld\n@@seconds %.6f\n", net->getNumFloatOps(), t1 - t0);
    print_plan(net);
    if (steps) print_steps(net);
    return 0;
}

// minfill qasm measure out.qbb reduce : in-process ordering, then contraction along it
static int mode_minfill(int argc, char **argv, bool steps) {
    auto net = std::make_shared<Network>(argv[2], argv[3]);
    if (atoi(argv[5])) net->ReduceCircuit();
    LineGraph lg(net);
    lg.SetQBBOutFiles("/dev/null", argv[4], "/dev/null");
    double t0 = now_s();
    const int width = lg.runMinFill();
    double t1 = now_s();
    bool ok = false;
    try { ok = lg.LGContract(); } catch (std::exception &e) { printf("@@exception %s\n", e.what()); }
    cplx v = net->GetFinalValue();
    printf("@@ok %d\n@@width %d\n@@ordering_seconds %.6f\n", (int)ok, width, t1 - t0);
    print_value("value", v);
    printf("@@flops %lld\n", net->getNumFloatOps());
    if (steps) print_steps(net);
    return 0;
}

static int mode_cnf(int argc, char **argv) {
    auto net = std::make_shared<Network>(argv[2], argv[3]);
    if (atoi(argv[5])) net->ReduceCircuit();
    LineGraph lg(net);
    std::ofstream out(argv[4]);
    lg.WriteCnf(out);
    printf("@@cnf %zu %zu\n", lg.NumLineGraphVertices(), lg.NumLineGraphEdges());
    return 0;
}

// inp script.inp: what the ".inp" reader makes of a script -- one line per map entry, sorted by key
static int mode_inp(int argc, char **argv) {
    (void)argc;
    leviParser script;
    const bool opened = script.readInputFile(argv[2]);
    printf("@@opened %d\n", opened ? 1 : 0);
    for (const auto &kv : script.mapString) printf("@@string %s %s\n", kv.first.c_str(), kv.second.c_str());
    for (const auto &kv : script.mapBool) printf("@@bool %s %d\n", kv.first.c_str(), kv.second ? 1 : 0);
    for (const auto &kv : script.mapInt) printf("@@int %s %d\n", kv.first.c_str(), kv.second);
    for (const auto &kv : script.mapDouble) printf("@@double %s %.17g\n", kv.first.c_str(), kv.second);
    return 0;
}

static int mode_seq(int argc, char **argv, bool steps) {
    const char *qasm = argv[2], *meas = argv[3], *planf = argv[4];
    std::vector<std::pair<int, int>> plan;
    FILE *f = fopen(planf, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", planf); return 2; }
    int a, b;
    while (fscanf(f, "%d %d", &a, &b) == 2) plan.push_back({a, b});
    fclose(f);
    ContractionTools tools(qasm, meas);
    double t0 = now_s();
    auto net = tools.ContractGivenSequence(plan);
    cplx v = tools.GetFinalVal();
    double t1 = now_s();
    printf("@@done %d\n", (int)net->IsDone());
    print_value("value", v);
    printf("@@flops %lld\n@@steps %zu\n@@seconds %.6f\n", net->getNumFloatOps(), plan.size(), t1 - t0);
    print_plan(net);
    if (steps) print_steps(net);
    return 0;
}

static int mode_stoch(int argc, char **argv, bool steps) {
    ContractionTools tools(argv[2], argv[3]);
    if (argc > 4) tools.SetSeed((unsigned)strtoul(argv[4], nullptr, 10));
    auto net = tools.Contract(Stochastic);
    print_value("value", tools.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    if (steps) print_steps(net);
    return 0;
}

// cost qasm measure pValue seed: ContractionTools::Contract(CostContractSimple, pValue) with a fixed seed
static int mode_cost(int argc, char **argv, bool steps) {
    ContractionTools tools(argv[2], argv[3]);
    tools.SetSeed((unsigned)strtoul(argv[5], nullptr, 10));
    auto net = tools.Contract(CostContractSimple, atoi(argv[4]));
    print_value("value", tools.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    if (steps) print_steps(net);
    return 0;
}

static int mode_user(int argc, char **argv) {
    ContractionTools tools(argv[2], argv[3]);
    auto net = tools.ContractUserDefinedSequenceOfWires(argv[4]);
    print_value("value", tools.GetFinalVal());
    printf("@@flops %lld\n", net->getNumFloatOps());
    print_plan(net);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: qtb_harness <mode> ...\n"); return 2; }
    std::string m = argv[1];
    bool steps = false;
    if (m.size() > 6 && m.substr(m.size() - 6) == "+steps") { steps = true; m = m.substr(0, m.size() - 6); }
    try {
        if (m == "gate") return mode_gate(argc, argv);
        if (m == "lg") return mode_lg(argc, argv, steps);
        if (m == "cnf") return mode_cnf(argc, argv);
        if (m == "inp") return mode_inp(argc, argv);
        if (m == "minfill") return mode_minfill(argc, argv, steps);
        if (m == "seq") return mode_seq(argc, argv, steps);
        if (m == "stoch") return mode_stoch(argc, argv, steps);
        if (m == "cost") return mode_cost(argc, argv, steps);
        if (m == "user") return mode_user(argc, argv);
    } catch (std::exception &e) {
        printf("@@exception %s\n", e.what());
        return 1;
    }
    fprintf(stderr, "unknown mode %s\n", m.c_str());
    return 2;
}
