// qtorch_b200/apps/host_capi.cpp -> qtorch_b200/libqtorch_host.so
// In-process C entry points over the C++ host mirror (Network / LineGraph / ContractionTools), so that
// bench.py and the tests can make "the call a user makes" -- build the Network from .qasm + measurement
// files, ReduceCircuit, contract along a frozen QuickBB ordering or a recorded plan, read the value --
// without paying process start-up and CUDA initialisation per call.  Same flow as qtorch's main.cpp
// (/root/reference/src/main.cpp:74-198).
#include <unistd.h>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../host/qtorch.hpp"
#include "../host/maxcut.h"

static thread_local std::string g_err;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" {

const char *qth_last_error(void) { return g_err.c_str(); }

// the process-wide engine context of the host mirror (for qtb_ctx_stats / timers / trace from the caller)
void *qth_engine_ctx(void) {
    try { return device::Engine::Get().ctx(); } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}

// Network(qasm, measure) -> [ReduceCircuit] -> LineGraph(net).LGContract() on the frozen ordering -> GetFinalValue().
// seconds_after_parse mirrors the reference's own stopwatch (started after the circuit is read, main.cpp:108-109).
int qth_contract_linegraph(const char *qasm, const char *measure, const char *qbbOut, int reduce, double value[2],
                           long long *flops, int *nodes, double *secondsAfterParse) {
    try {
        auto net = std::make_shared<Network>(qasm, measure);
        const double t0 = now_s();
        if (reduce) net->ReduceCircuit();
        LineGraph lg(net);
        lg.SetQBBOutFiles("/dev/null", qbbOut, "/dev/null");
        const bool ok = lg.LGContract();
        const std::complex<double> v = net->GetFinalValue();
        if (secondsAfterParse) *secondsAfterParse = now_s() - t0;
        value[0] = v.real(); value[1] = v.imag();
        if (flops) *flops = net->getNumFloatOps();
        if (nodes) *nodes = static_cast<int>(net->GetAllNodes().size());
        return ok ? 0 : 2;
    } catch (std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// The same value through the plan cache (host/PlanCache.h): the first call for a (qasm, ordering, reduce) triple does the
// host bookkeeping once and compiles the plan; every later call only swaps the measurement caps and replays the graph.
int qth_contract_cached(const char *qasm, const char *measure, const char *qbbOut, int reduce, double value[2], long long *flops, int *nodes,
                        int *cacheHit) {
    try {
        PlanCache &cache = PlanCache::Get();
        const long long hitsBefore = cache.Hits();
        std::shared_ptr<CompiledCircuit> c = cache.Lookup(qasm, qbbOut ? qbbOut : "", reduce != 0);
        if (cacheHit) *cacheHit = cache.Hits() > hitsBefore ? 1 : 0;
        const std::complex<double> v = c->EvaluateFile(measure);
        value[0] = v.real(); value[1] = v.imag();
        if (flops) *flops = c->Units();
        if (nodes) *nodes = c->NumNodes();
        return c->Ok() ? 0 : 2;
    } catch (std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
void qth_plan_cache_clear(void) { PlanCache::Get().Clear(); }

// The same call split in two, so that a caller with many networks (the 60 <ZiZj> terms of one QAOA evaluation) can
// overlap the host bookkeeping of network i+1 with the device work of network i: `begin` parses, reduces, walks the
// ordering and enqueues every step (no synchronisation -- Network::GetFinalValue is lazy); `end` reads the scalar back.
struct QthJob { std::shared_ptr<Network> net; bool ok; };

void *qth_linegraph_begin(const char *qasm, const char *measure, const char *qbbOut, int reduce) {
    try {
        auto net = std::make_shared<Network>(qasm, measure);
        if (reduce) net->ReduceCircuit();
        LineGraph lg(net);
        lg.SetQBBOutFiles("/dev/null", qbbOut, "/dev/null");
        const bool ok = lg.LGContract();
        net->PrefetchFinalValue();        // launches what is still held back and queues the scalar read right behind it
        return new QthJob{net, ok};
    } catch (std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

int qth_linegraph_end(void *job, double value[2], long long *flops, int *nodes) {
    if (!job) { g_err = "null job"; return 1; }
    QthJob *j = static_cast<QthJob *>(job);
    int rc = j->ok ? 0 : 2;
    try {
        const std::complex<double> v = j->net->GetFinalValue();
        value[0] = v.real(); value[1] = v.imag();
        if (flops) *flops = j->net->getNumFloatOps();
        if (nodes) *nodes = static_cast<int>(j->net->GetAllNodes().size());
    } catch (std::exception &e) {
        g_err = e.what();
        rc = 1;
    }
    delete j;
    return rc;
}

// ContractionTools(qasm, measure).ContractGivenSequence(pairs): replay of a recorded plan (mCreatedFrom pairs).
int qth_contract_sequence(const char *qasm, const char *measure, const int *pairs, int nPairs, double value[2],
                          long long *flops, int *nodes, double *secondsAfterParse) {
    try {
        std::vector<std::pair<int, int>> seq(nPairs);
        for (int i = 0; i < nPairs; i++) seq[i] = {pairs[2 * i], pairs[2 * i + 1]};
        auto net = std::make_shared<Network>(qasm, measure);
        ContractionTools tools(net);
        const double t0 = now_s();
        tools.ContractGivenSequence(seq);
        const std::complex<double> v = tools.GetFinalVal();
        if (secondsAfterParse) *secondsAfterParse = now_s() - t0;
        value[0] = v.real(); value[1] = v.imag();
        if (flops) *flops = net->getNumFloatOps();
        if (nodes) *nodes = static_cast<int>(net->GetAllNodes().size());
        return 0;
    } catch (std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// Plan export without touching the device (host bookkeeping only): runs the same flow in plan-only mode and
// returns the step list in qtb_plan_step layout plus every original node's tensor (the plan's inputs).
// Buffers are owned by the library until the next call on this thread.
struct QthPlan {
    int nInputs, nSteps;
    const int *inputRanks;
    const double *inputData;        // concatenated (re,im) of all inputs in id order
    const long long *inputOffsets;  // element offset (in complex numbers) of input i inside inputData
    const qtb_plan_step *steps;
    long long flops;
};
static thread_local std::vector<int> t_ranks;
static thread_local std::vector<double> t_data;
static thread_local std::vector<long long> t_offs;
static thread_local std::vector<qtb_plan_step> t_steps;

int qth_export_plan_linegraph(const char *qasm, const char *measure, const char *qbbOut, int reduce, QthPlan *out) {
    const bool before = device::Engine::PlanOnly();
    device::Engine::SetPlanOnly(true);
    int rc = 0;
    try {
        auto net = std::make_shared<Network>(qasm, measure);
        const int n = static_cast<int>(net->GetAllNodes().size());
        t_ranks.clear(); t_data.clear(); t_offs.clear(); t_steps.clear();
        for (int i = 0; i < n; i++) {
            auto &node = net->GetAllNodes()[i];
            t_ranks.push_back(node->mRank);
            t_offs.push_back(static_cast<long long>(t_data.size() / 2));
            for (size_t e = 0; e < node->NumElements(); e++) {
                const std::complex<double> &v = node->Access(static_cast<long long>(e));
                t_data.push_back(v.real()); t_data.push_back(v.imag());
            }
        }
        if (reduce) net->ReduceCircuit();
        LineGraph lg(net);
        if (qbbOut && qbbOut[0]) {
            lg.SetQBBOutFiles("/dev/null", qbbOut, "/dev/null");
        } else {
            // no frozen QuickBB file: in-process min-fill ordering (LineGraph::runMinFill)
            lg.SetQBBOutFiles("/dev/null", "", "/dev/null");          // ordering handed over in memory, no temp file
            lg.runMinFill();
        }
        if (!lg.LGContract()) rc = 2;
        for (const auto &r : net->GetPlan()) {
            qtb_plan_step s;
            memset(&s, 0, sizeof(s));
            s.a = r.a; s.b = r.b; s.k = static_cast<int>(r.posA.size());
            for (int j = 0; j < s.k; j++) { s.pos_a[j] = static_cast<int8_t>(r.posA[j]); s.pos_b[j] = static_cast<int8_t>(r.posB[j]); }
            t_steps.push_back(s);
        }
        out->nInputs = n; out->nSteps = static_cast<int>(t_steps.size());
        out->inputRanks = t_ranks.data(); out->inputData = t_data.data(); out->inputOffsets = t_offs.data();
        out->steps = t_steps.data(); out->flops = net->getNumFloatOps();
    } catch (std::exception &e) {
        g_err = e.what();
        rc = 1;
    }
    device::Engine::SetPlanOnly(before);
    return rc;
}

// ---- index slicing (host/Slicing.h) ------------------------------------------------------------------------------------
// planner only (no device): cut `nWires` wires of a plan given in qtb_plan_step layout.  Buffers are owned by the library
// until the next call on this thread.  wires / cuts use the label encoding of Slicing.h (input tensor * 32 + leg).
struct QthSlicedPlan {
    int nInputs, nSteps, nInvariant, nWires, peakRank, nCuts;
    const int *inputRanks;          // ranks of ONE slice's inputs
    const qtb_plan_step *steps;     // hoisted sliced plan
    const int *wires;               // nWires labels
    const int *cuts;                // nCuts triples (input tensor, leg, wire label)
    double unitsPerSlice, unitsInvariant;
};
static thread_local std::vector<int> t_sRanks, t_sWires, t_sCuts;
static thread_local std::vector<qtb_plan_step> t_sSteps;

static slicing::Plan planFromAbi(int nInputs, const int *ranks, int nSteps, const qtb_plan_step *steps) {
    slicing::Plan p;
    p.inputRanks.assign(ranks, ranks + nInputs);
    for (int i = 0; i < nSteps; i++) {
        slicing::Step s{steps[i].a, steps[i].b, {}, {}};
        for (int j = 0; j < steps[i].k; j++) { s.posA.push_back(steps[i].pos_a[j]); s.posB.push_back(steps[i].pos_b[j]); }
        p.steps.push_back(s);
    }
    return p;
}

int qth_slice_plan(int nInputs, const int *ranks, int nSteps, const qtb_plan_step *steps, int nWires, QthSlicedPlan *out) {
    try {
        const slicing::Plan plan = planFromAbi(nInputs, ranks, nSteps, steps);
        const slicing::SlicedPlan sp = slicing::SlicePlan(plan, slicing::ChooseWires(plan, nWires));
        t_sRanks = sp.plan.inputRanks;
        t_sSteps = slicing::ToAbiSteps(sp.plan);
        t_sWires = sp.wires;
        t_sCuts.clear();
        for (const auto &c : sp.cuts) for (const auto &lw : c.second) { t_sCuts.push_back(c.first); t_sCuts.push_back(lw.first); t_sCuts.push_back(lw.second); }
        out->nInputs = nInputs; out->nSteps = static_cast<int>(t_sSteps.size()); out->nInvariant = sp.nInvariant;
        out->nWires = static_cast<int>(sp.wires.size()); out->peakRank = sp.peakRank; out->nCuts = static_cast<int>(t_sCuts.size() / 3);
        out->inputRanks = t_sRanks.data(); out->steps = t_sSteps.data(); out->wires = t_sWires.data(); out->cuts = t_sCuts.data();
        out->unitsPerSlice = static_cast<double>(sp.unitsPerSlice); out->unitsInvariant = static_cast<double>(sp.unitsInvariant);
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return 1; }
}

// one input tensor's slice (host only): legs[i] fixed to digits[i]
int qth_slice_tensor(const double *full, int rank, const int *legs, const int *digits, int nCut, double *out) {
    try {
        std::vector<std::complex<double>> f(static_cast<size_t>(1) << (2 * rank));
        memcpy(f.data(), full, f.size() * 16);
        std::vector<std::pair<int, int>> cut;
        for (int i = 0; i < nCut; i++) cut.push_back({legs[i], digits[i]});
        const auto r = slicing::SliceTensor(f, rank, cut);
        memcpy(out, r.data(), r.size() * 16);
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return 1; }
}

// A line-graph network, index-sliced over the ranks of the job and run by the sliced-amplitude executor.  The host
// bookkeeping (parse, ReduceCircuit, ordering walk) runs once in plan-only mode; the caller's communicator must already be
// initialised on the host mirror's engine context when world > 1 (qtb_comm_init on qth_engine_ctx()).
struct QthSliced {
    std::unique_ptr<SlicedContraction> run;
    SlicedContraction::Tensors inputs;
    long long flopsUnsliced = 0;
};

void *qth_sliced_create(const char *qasm, const char *measure, const char *qbbOut, int reduce, int nSliceWires, int lanes, int rank, int world) {
    const bool before = device::Engine::PlanOnly();
    device::Engine::SetPlanOnly(true);
    QthSliced *h = new QthSliced();
    slicing::Plan plan;
    try {
        auto net = std::make_shared<Network>(qasm, measure);
        for (int i = 0; i < net->GetNumOriginalNodes(); i++) h->inputs.push_back(net->GetAllNodes()[i]->GetTensorVals());
        if (reduce) net->ReduceCircuit();
        LineGraph lg(net);
        if (qbbOut && qbbOut[0]) lg.SetQBBOutFiles("/dev/null", qbbOut, "/dev/null");
        else { lg.SetQBBOutFiles("/dev/null", "", "/dev/null"); lg.runMinFill(); }
        lg.LGContract();
        plan = slicing::PlanOfNetwork(*net);
        h->flopsUnsliced = net->getNumFloatOps();
    } catch (std::exception &e) {
        g_err = e.what();
        device::Engine::SetPlanOnly(before);
        delete h;
        return nullptr;
    }
    device::Engine::SetPlanOnly(before);
    try {
        // world < 0: this process takes rank's share of a |world|-rank job but skips the reduction (profiling one rank's load)
        device::Job job;
        job.rank = rank; job.world = world < 0 ? -world : world;
        h->run.reset(new SlicedContraction(plan, nSliceWires, job, lanes, world > 0));
    } catch (std::exception &e) { g_err = e.what(); delete h; return nullptr; }
    return h;
}
void qth_sliced_destroy(void *h) { delete static_cast<QthSliced *>(h); }
// info[0..7]: slices, owned, invariant steps, steps, peak rank of a slice, cut wires, launches per slice, launches of the prefix
int qth_sliced_info(void *hh, long long *info, double *unitsPerSlice, double *unitsInvariant, long long *flopsUnsliced) {
    QthSliced *h = static_cast<QthSliced *>(hh);
    const slicing::SlicedPlan &sp = h->run->Sliced();
    info[0] = static_cast<long long>(sp.NumSlices()); info[1] = static_cast<long long>(h->run->OwnedSlices().size());
    info[2] = sp.nInvariant; info[3] = static_cast<long long>(sp.plan.steps.size()); info[4] = sp.peakRank; info[5] = static_cast<long long>(sp.wires.size());
    int pre = 0;
    info[6] = h->run->LaunchesPerSlice(&pre); info[7] = pre;
    *unitsPerSlice = static_cast<double>(sp.unitsPerSlice); *unitsInvariant = static_cast<double>(sp.unitsInvariant);
    *flopsUnsliced = h->flopsUnsliced;
    return 0;
}
int qth_sliced_stage(void *hh, int bank) {
    QthSliced *h = static_cast<QthSliced *>(hh);
    try { h->run->Stage(h->inputs, bank); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int qth_sliced_begin(void *hh, int bank) {
    QthSliced *h = static_cast<QthSliced *>(hh);
    try { return h->run->Begin(bank); } catch (std::exception &e) { g_err = e.what(); return -1; }
}
int qth_sliced_end(void *hh, int ticket, double value[2]) {
    QthSliced *h = static_cast<QthSliced *>(hh);
    try { const std::complex<double> v = h->run->End(ticket); value[0] = v.real(); value[1] = v.imag(); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}

// ---- QAOA term dispatcher (host/maxcut.h QaoaObjective) -------------------------------------------------------------
void *qth_qaoa_create(const char *graphFile, int p, int rank, int world, int planTries) {
    try {
        ExtraData data(p, graphFile);
        return new QaoaObjective(data, rank, world, nullptr, planTries);
    } catch (std::exception &e) { g_err = e.what(); return nullptr; } catch (const char *m) { g_err = m; return nullptr; }
}
void qth_qaoa_destroy(void *h) { delete static_cast<QaoaObjective *>(h); }
int qth_qaoa_num_owned(void *h) { return static_cast<int>(static_cast<QaoaObjective *>(h)->OwnedEdges().size()); }
int qth_qaoa_owned_edges(void *h, int *out) {
    const auto &o = static_cast<QaoaObjective *>(h)->OwnedEdges();
    for (size_t i = 0; i < o.size(); i++) out[i] = o[i];
    return static_cast<int>(o.size());
}
long long qth_qaoa_units(void *h) { return static_cast<QaoaObjective *>(h)->UnitsPerEvaluation(); }
int qth_qaoa_launches(void *h) { return static_cast<QaoaObjective *>(h)->LaunchesPerEvaluation(); }
// <ZZ> of every owned edge (re, im pairs) and this rank's partial F_p for the angles (beta_1..p, gamma_1..p)
int qth_qaoa_evaluate(void *h, const double *betasGammas, int n, double *termsReIm, double *partialFp) {
    try {
        std::vector<double> bg(betasGammas, betasGammas + n);
        const auto vals = static_cast<QaoaObjective *>(h)->EvaluateOwnedTerms(bg);
        double fp = 0.0;
        for (size_t i = 0; i < vals.size(); i++) {
            if (termsReIm) { termsReIm[2 * i] = vals[i].real(); termsReIm[2 * i + 1] = vals[i].imag(); }
            fp += 0.5 * (1.0 - vals[i].real());
        }
        if (partialFp) *partialFp = fp;
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// F_p for the angles through QaoaObjective's own path: one graph launch; with reduce != 0 the sum over ranks is ONE
// in-stream ncclAllReduce (the communicator must be initialised on qth_engine_ctx()) and every rank gets the full F_p
int qth_qaoa_objective(void *h, const double *betasGammas, int n, int reduce, int numEdgesTotal, double *fp) {
    try {
        QaoaObjective *q = static_cast<QaoaObjective *>(h);
        std::vector<double> bg(betasGammas, betasGammas + n);
        q->Begin(bg, reduce != 0);
        const std::complex<double> sum = q->End();
        const double nEdges = reduce ? static_cast<double>(numEdgesTotal) : static_cast<double>(q->OwnedEdges().size());
        *fp = 0.5 * (nEdges - sum.real());
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// host-only (no device): light-cone circuit text + measurement string of one edge, straight from ExtraData
int qth_maxcut_circuit_text(const char *graphFile, int p, int edge, const double *betasGammas, char *buf, int bufLen, int *numEdges, int *numQubits) {
    try {
        ExtraData data(p, graphFile);
        if (numEdges) *numEdges = static_cast<int>(data.pairs.size());
        if (edge < 0 || edge >= static_cast<int>(data.pairs.size())) return -1;
        std::vector<double> bg(betasGammas, betasGammas + 2 * p);
        std::ostringstream q;
        const int nq = data.qubitsNeeded[edge];
        if (numQubits) *numQubits = nq;
        q << nq << std::endl;
        outputInitialPlusStateToFile(q, nq);
        applyU_CsThenU_Bs(data.realIterations[edge], data.p, bg, nq, q);
        const std::string t = q.str();
        if (buf && bufLen > 0) { strncpy(buf, t.c_str(), bufLen - 1); buf[bufLen - 1] = 0; }
        return static_cast<int>(t.size());
    } catch (std::exception &e) { g_err = e.what(); return -2; } catch (const char *m) { g_err = m; return -2; }
}
// circuit text the reference would write to input/tempMaxCut.qasm for this edge; returns the length
int qth_qaoa_circuit_text(void *h, int edge, const double *betasGammas, int n, char *buf, int bufLen) {
    std::vector<double> bg(betasGammas, betasGammas + n);
    const std::string t = static_cast<QaoaObjective *>(h)->CircuitText(edge, bg);
    if (buf && bufLen > 0) { strncpy(buf, t.c_str(), bufLen - 1); buf[bufLen - 1] = 0; }
    return static_cast<int>(t.size());
}

// final cut string for given angles; bits[] receives numQubits 0/1 values; returns the number of qubits or -1
int qth_maxcut_final_string(const char *graphFile, int p, const double *betasGammas, const char *outFile, int *bits, int maxBits, unsigned seed,
                            double *probability) {
    try {
        std::vector<double> bg(betasGammas, betasGammas + 2 * p);
        const std::vector<bool> ans = maxcutGetFinalString(graphFile, p, {}, bg, outFile, seed, probability);
        for (size_t i = 0; i < ans.size() && static_cast<int>(i) < maxBits; i++) bits[i] = ans[i] ? 1 : 0;
        return static_cast<int>(ans.size());
    } catch (std::exception &e) { g_err = e.what(); return -1; } catch (const char *m) { g_err = m; return -1; }
}

}  // extern "C"
