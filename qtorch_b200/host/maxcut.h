// qtorch_b200/host/maxcut.h -- QAOA MaxCut helpers (counterpart of /root/reference/src/maxcut.h:31-233) and the
// B200 term dispatcher that replaces the serial per-edge loop of F_p (/root/reference/src/maxcut.cpp:162-204).
//
// Kept from the reference (same names, same results): ExtraData (graph reader + per-edge light-cone extraction with
// its qubit-relabelling order), outputInitialPlusStateToFile, applyU_CsThenU_Bs (now writing to any std::ostream).
// New: QaoaObjective -- every edge's light-cone circuit is built IN MEMORY (no input/tempMaxCut.qasm round trip;
// angles still pass through the 6-significant-digit text form and std::stof, so tensors are bit-identical to what
// the reference parses back), planned once (plan topology does not depend on the angles), compiled to a device
// plan, and all edges owned by this rank are evaluated as ONE CUDA-graph launch per objective evaluation (qtb_batch_*:
// only the 2p gate tables travel to the device, the plans' other inputs stay resident).  Edges are dealt round-robin over
// ranks; the sum meets in one in-stream NCCL allreduce.
#pragma once

#include <fstream>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ContractionTools.h"
#include "preprocess.h"

namespace qtorch {

struct ExtraData {
    ExtraData(const int p0, const char *filename0) : fileName(filename0), p(p0) { ReadInData(); PopulateIterations(); }
    ExtraData() {}
    std::string fileName;
    std::vector<std::pair<int, int>> pairs;                      // graph edges in file order
    std::vector<std::vector<int>> adjacencyLists;
    std::vector<int> qubitsNeeded;                               // light-cone size per edge
    std::string outputFile;
    int numQubits{0};
    int p{1};
    std::vector<std::vector<std::pair<int, int>>> iterations;     // light-cone edges per term, graph vertex ids
    std::vector<std::vector<std::pair<int, int>>> realIterations; // the same in the term's own qubit numbering

    // ".dgf": "c ..." comment lines, "e u v" edges; vertex count = largest id + 1 (reference maxcut.h:56-103)
    void ReadInData() {
        std::ifstream input(fileName);
        if (!input.is_open()) {
            std::cout << "Could Not Open File" << std::endl;
            throw "File Not Open";
        }
        int largest = 0;
        char tag;
        std::string rest;
        while (input >> tag) {
            if (tag != 'e') {
                if (tag != 'c') std::cout << "Error parsing file" << std::endl;
                std::getline(input, rest);
                continue;
            }
            int u, v;
            input >> u >> v;
            largest = std::max(largest, std::max(u, v));
            pairs.push_back({u, v});
        }
        numQubits = largest + 1;
        adjacencyLists.assign(numQubits, std::vector<int>());
        for (const auto &e : pairs) {
            adjacencyLists[e.first].push_back(e.second);
            adjacencyLists[e.second].push_back(e.first);
        }
    }

    // For every edge: breadth-first growth of the light cone, p rounds.  Term-local qubit ids are handed out in
    // order of first appearance while scanning adjacency lists -- so qubit 1 is the FIRST LISTED NEIGHBOUR of the
    // edge's first vertex, not necessarily its partner (reference quirk, maxcut.h:105-185).
    void PopulateIterations() {
        const size_t nTerms = pairs.size();
        qubitsNeeded.assign(nTerms, 0);
        iterations.assign(nTerms, std::vector<std::pair<int, int>>());
        realIterations.assign(nTerms, std::vector<std::pair<int, int>>());
        std::vector<int> localId(numQubits);
        for (size_t term = 0; term < nTerms; ++term) {
            std::fill(localId.begin(), localId.end(), -1);
            std::vector<int> frontier{pairs[term].first, pairs[term].second};
            localId[pairs[term].first] = 0;
            qubitsNeeded[term] = 1;
            std::vector<bool> next(numQubits, false), done(numQubits, false), inFrontier(numQubits, false);
            inFrontier[pairs[term].first] = inFrontier[pairs[term].second] = true;
            PopulateIterationsHelper(0, frontier, next, done, inFrontier, static_cast<int>(term), localId);
            for (const auto &e : iterations[term]) realIterations[term].push_back({localId[e.first], localId[e.second]});
        }
    }

    void PopulateIterationsHelper(int counter, std::vector<int> &workingVerticesList, std::vector<bool> &newWorkingVertices,
                                  std::vector<bool> &hasBeenChecked, std::vector<bool> &isInWorkingNodes, int iterationIndex,
                                  std::vector<int> &mapToRealIt) {
        for (int round = counter; round < p; ++round) {
            for (int v : workingVerticesList) {
                for (int nb : adjacencyLists[v]) {
                    if (hasBeenChecked[nb]) continue;
                    if (mapToRealIt[nb] == -1) mapToRealIt[nb] = qubitsNeeded[iterationIndex]++;
                    iterations[iterationIndex].push_back({v, nb});
                    if (!isInWorkingNodes[nb]) newWorkingVertices[nb] = true;
                }
                hasBeenChecked[v] = true;
            }
            workingVerticesList.clear();
            std::fill(isInWorkingNodes.begin(), isInWorkingNodes.end(), false);
            for (int v = 0; v < static_cast<int>(newWorkingVertices.size()); ++v) {
                if (newWorkingVertices[v]) {
                    workingVerticesList.push_back(v);
                    isInWorkingNodes[v] = true;
                }
            }
            std::fill(newWorkingVertices.begin(), newWorkingVertices.end(), false);
        }
    }
};

// |+>^n
inline void outputInitialPlusStateToFile(std::ostream &qasm, const int numQubits) {
    for (int q = 0; q < numQubits; ++q) qasm << "H " << q << std::endl;
}

// p layers of  [ CNOT a b ; Rz(-gamma) b ; CNOT a b  per light-cone edge ]  then  Rx(2 beta) on every qubit.
// betas_gammas = (beta_1..beta_p, gamma_1..gamma_p); default stream precision (6 significant digits), as the reference.
inline void applyU_CsThenU_Bs(const std::vector<std::pair<int, int>> &objectiveF, const int p, const std::vector<double> &betas_gammas,
                              const int numQubits, std::ostream &output) {
    for (int layer = 0; layer < p; ++layer) {
        const double gamma = betas_gammas[layer + p], beta = betas_gammas[layer];
        for (const auto &e : objectiveF) {
            output << "CNOT " << e.first << " " << e.second << std::endl;
            output << "Rz " << -gamma << " " << e.second << std::endl;
            output << "CNOT " << e.first << " " << e.second << std::endl;
        }
        for (int q = 0; q < numQubits; ++q) output << "Rx " << beta * 2.0 << " " << q << std::endl;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// The B200 term dispatcher
class QaoaObjective {
public:
    // rank/world: which share of the edges this process owns (edge e belongs to rank e % world).
    // allreduce: non-null when the process belongs to a multi-rank job (device::Job::FromEnvironment has joined the NCCL
    // communicator on the engine context): operator() then sums the objective over all ranks with ONE in-stream
    // ncclAllReduce per evaluation.  The functor itself is only used by callers that reduce on the host.
    QaoaObjective(const ExtraData &data, int rank = 0, int world = 1, std::function<void(double *, int)> allreduce = nullptr,
                  int planTries = 8)
        : mData(data), mRank(rank), mWorld(world), mAllReduce(allreduce) {
        for (size_t e = static_cast<size_t>(rank); e < mData.pairs.size(); e += static_cast<size_t>(world)) mOwned.push_back(static_cast<int>(e));
        std::vector<double> bg(2 * mData.p);
        for (int i = 0; i < mData.p; ++i) { bg[i] = 0.392699; bg[i + mData.p] = 0.785399; }      // maxcut.cpp:155-157
        // many small plans evaluated side by side (one CTA each): let larger steps ride in the grouped launch, so that a
        // whole evaluation is ONE launch of grouped micro-steps between the table scatter and the gather
        qtb_ctx *ctx = device::Engine::Get().ctx();
        const int before = qtb_ctx_get_micro_limit(ctx);
        device::check(qtb_ctx_set_micro_limit(ctx, 10));
        try {
            for (int e : mOwned) mTerms.push_back(BuildTerm(e, bg, planTries));
        } catch (...) {
            qtb_ctx_set_micro_limit(ctx, before);
            DestroyPlans();
            throw;
        }
        device::check(qtb_ctx_set_micro_limit(ctx, before));
        if (mTerms.empty()) return;
        // the batch: plans resident, table l = Rz(-gamma_l), table p + l = Rx(2 beta_l)
        std::vector<qtb_plan *> plans;
        for (auto &t : mTerms) plans.push_back(t.plan);
        try {
            device::check(qtb_batch_create(ctx, plans.data(), static_cast<int>(plans.size()), 2 * mData.p, 2, &mBatch));
            for (size_t i = 0; i < mTerms.size(); ++i) {
                device::check(qtb_batch_set_inputs(ctx, mBatch, static_cast<int>(i), mTerms[i].ptrs.data()));
                for (const auto &sl : mTerms[i].slots)
                    device::check(qtb_batch_bind(ctx, mBatch, static_cast<int>(i), sl.input, sl.isRx ? mData.p + sl.layer : sl.layer));
            }
        } catch (...) {
            DestroyPlans();
            throw;
        }
    }
    ~QaoaObjective() { DestroyPlans(); }
    QaoaObjective(const QaoaObjective &) = delete;
    QaoaObjective &operator=(const QaoaObjective &) = delete;

    // circuit text of one edge's light cone for the given angles (what the reference writes to input/tempMaxCut.qasm)
    std::string CircuitText(int edge, const std::vector<double> &betas_gammas) const {
        std::ostringstream q;
        const int nq = mData.qubitsNeeded[edge];
        q << nq << std::endl;
        outputInitialPlusStateToFile(q, nq);
        applyU_CsThenU_Bs(mData.realIterations[edge], mData.p, betas_gammas, nq, q);
        return q.str();
    }
    static std::string MeasurementText(int numQubits) {
        std::string m;
        for (int i = 0; i < numQubits; ++i) m += (i < 2) ? "Z " : "T ";
        return m;
    }

    // start one evaluation: 2p gate tables up, one graph launch, [one in-stream allreduce of the sum], results queued
    void Begin(const std::vector<double> &betas_gammas, bool reduceOverRanks) {
        if (mTerms.empty() && !(reduceOverRanks && mWorld > 1)) return;
        if (mTerms.empty()) throw InvalidFunctionInput();             // a rank without edges cannot join the reduction (more ranks than edges)
        const std::vector<std::complex<double>> tables = GateTables(betas_gammas);
        device::check(qtb_batch_begin(device::Engine::Get().ctx(), mBatch, reinterpret_cast<const double *>(tables.data()), reduceOverRanks && mWorld > 1 ? 1 : 0));
    }
    // wait for it: sum of <Z Z> over this rank's edges (over all edges of all ranks after a reduction); `terms`, if given,
    // receives this rank's individual values
    std::complex<double> End(std::vector<std::complex<double>> *terms = nullptr) {
        double sum[2] = {0.0, 0.0};
        if (terms) terms->assign(mTerms.size(), std::complex<double>(0.0));
        if (!mTerms.empty())
            device::check(qtb_batch_end(device::Engine::Get().ctx(), mBatch, sum, terms ? reinterpret_cast<double *>(terms->data()) : nullptr));
        return {sum[0], sum[1]};
    }

    // <Z Z> of every edge owned by this rank, one graph launch
    std::vector<std::complex<double>> EvaluateOwnedTerms(const std::vector<double> &betas_gammas) {
        std::vector<std::complex<double>> vals;
        Begin(betas_gammas, false);
        End(&vals);
        return vals;
    }

    // F_p = sum_edges 1/2 (1 - Re<Z Z>)   (maxcut.cpp:196), summed over ranks
    double operator()(const std::vector<double> &betas_gammas) {
        const bool reduce = mAllReduce && mWorld > 1;
        Begin(betas_gammas, reduce);
        const std::complex<double> sum = End();
        ++mEvaluations;
        const double nEdges = reduce ? static_cast<double>(mData.pairs.size()) : static_cast<double>(mOwned.size());
        return 0.5 * (nEdges - sum.real());
    }

    const std::vector<int> &OwnedEdges() const { return mOwned; }
    long long Evaluations() const { return mEvaluations; }
    long long UnitsPerEvaluation() const { long long u = 0; for (const auto &t : mTerms) u += t.units; return u; }
    int LaunchesPerEvaluation() const { return mBatch ? qtb_batch_launches(mBatch) : 0; }

private:
    struct AngleSlot { int input; bool isRx; int layer; };     // which plan input is Rz(-gamma_layer) / Rx(2 beta_layer)
    struct Term {
        int edge{0};
        qtb_plan *plan{nullptr};
        long long units{0};
        std::vector<std::vector<std::complex<double>>> inputs;   // host tensors of all original nodes, id order
        std::vector<const double *> ptrs;
        std::vector<AngleSlot> slots;
    };

    void DestroyPlans() {
        if (!device::Engine::Get().alive()) return;
        qtb_ctx *ctx = device::Engine::Get().ctx();
        if (mBatch) { qtb_batch_destroy(ctx, mBatch); mBatch = nullptr; }
        for (auto &t : mTerms) if (t.plan) { qtb_plan_destroy(ctx, t.plan); t.plan = nullptr; }
    }

    // the value the reference's parser would obtain: default-precision text, then std::stof (Network.h:342,402)
    static double ThroughText(double angle) {
        std::ostringstream os;
        os << angle;
        return static_cast<double>(std::stof(os.str()));
    }

    // the 2p gate tensors of one evaluation, built by the same constructors the parser uses: [Rz(-gamma_l)]_l, [Rx(2 beta_l)]_l
    std::vector<std::complex<double>> GateTables(const std::vector<double> &bg) const {
        std::vector<std::complex<double>> tables;
        for (int l = 0; l < mData.p; ++l) {
            RzNode g(ThroughText(-bg[l + mData.p]));
            tables.insert(tables.end(), g.GetTensorVals().begin(), g.GetTensorVals().end());
        }
        for (int l = 0; l < mData.p; ++l) {
            RxNode g(ThroughText(bg[l] * 2.0));
            tables.insert(tables.end(), g.GetTensorVals().begin(), g.GetTensorVals().end());
        }
        return tables;
    }

    Term BuildTerm(int edge, const std::vector<double> &bg, int planTries) {
        Term t;
        t.edge = edge;
        const int nq = mData.qubitsNeeded[edge];
        const std::string text = CircuitText(edge, bg), meas = MeasurementText(nq);
        // plan search on the host only (plan-only mode records steps without arithmetic): a few seeded stochastic searches
        // (the reference's own planner for this path, maxcut.cpp:189-190) and the in-process min-fill line-graph order;
        // cheapest by the reference's own unit count.  The plan does not depend on the angles.
        const bool before = device::Engine::PlanOnly();
        device::Engine::SetPlanOnly(true);
        std::shared_ptr<Network> best;
        try {
            for (int attempt = 0; attempt <= planTries; ++attempt) {
                std::istringstream qs(text);
                std::shared_ptr<Network> net = std::make_shared<Network>(qs, meas);
                if (attempt == 0) SnapshotInputs(*net, t);
                if (attempt < planTries) {
                    ContractionTools tools(net);
                    tools.SetSeed(1000003u * static_cast<unsigned>(edge) + static_cast<unsigned>(attempt));
                    tools.Contract(Stochastic);
                } else {
                    net->ReduceCircuit();
                    LineGraph lg(net);
                    lg.SetQBBOutFiles("/dev/null", "", "/dev/null");
                    lg.runMinFill();
                    lg.LGContract();
                }
                if (std::getenv("QTB_QAOA_VERBOSE")) {
                    std::vector<int> lv(net->GetNumOriginalNodes(), 0);
                    int depth = 0;
                    for (const auto &r : net->GetPlan()) { const int l = 1 + std::max(lv[r.a], lv[r.b]); lv.push_back(l); depth = std::max(depth, l); }
                    std::cerr << "edge " << edge << " attempt " << attempt << (attempt < planTries ? " stochastic" : " minfill") << ": steps " << net->GetPlan().size()
                              << " levels " << depth << " units " << net->getNumFloatOps() << std::endl;
                }
                if (!best || net->getNumFloatOps() < best->getNumFloatOps()) best = net;
            }
        } catch (...) {
            device::Engine::SetPlanOnly(before);
            throw;
        }
        device::Engine::SetPlanOnly(before);
        t.units = best->getNumFloatOps();
        std::vector<qtb_plan_step> steps;
        for (const auto &r : best->GetPlan()) {
            qtb_plan_step s;
            std::memset(&s, 0, sizeof(s));
            s.a = r.a; s.b = r.b; s.k = static_cast<int>(r.posA.size());
            for (int j = 0; j < s.k; ++j) { s.pos_a[j] = static_cast<int8_t>(r.posA[j]); s.pos_b[j] = static_cast<int8_t>(r.posB[j]); }
            steps.push_back(s);
        }
        std::vector<int> ranks;
        for (const auto &in : t.inputs) { int r = 0; while ((static_cast<size_t>(1) << (2 * r)) < in.size()) ++r; ranks.push_back(r); }
        device::check(qtb_plan_create(device::Engine::Get().ctx(), static_cast<int>(ranks.size()), ranks.data(), static_cast<int>(steps.size()),
                                      steps.data(), &t.plan));
        for (const auto &in : t.inputs) t.ptrs.push_back(reinterpret_cast<const double *>(in.data()));
        return t;
    }

    // copy every original node's tensor; remember which ones carry an angle (file order: per layer, per light-cone
    // edge one Rz, then one Rx per qubit)
    void SnapshotInputs(Network &net, Term &t) {
        const int n = net.GetNumOriginalNodes();
        const int rzPerLayer = static_cast<int>(mData.realIterations[t.edge].size()), rxPerLayer = mData.qubitsNeeded[t.edge];
        int rzCount = 0, rxCount = 0;
        for (int i = 0; i < n; ++i) {
            const std::shared_ptr<Node> &node = net.GetAllNodes()[i];
            t.inputs.push_back(node->GetTensorVals());
            if (node->GetTypeOfNode() == GateType::RZ) t.slots.push_back({i, false, rzCount++ / rzPerLayer});
            else if (node->GetTypeOfNode() == GateType::RX) t.slots.push_back({i, true, rxCount++ / rxPerLayer});
        }
    }

    ExtraData mData;
    int mRank, mWorld;
    std::function<void(double *, int)> mAllReduce;
    std::vector<int> mOwned;
    std::vector<Term> mTerms;
    qtb_batch *mBatch{nullptr};
    long long mEvaluations{0};
};

// ----------------------------------------------------------------------------------------------------------------
// Final cut string (counterpart of maxcutGetFinalString, /root/reference/src/maxcut.cpp:29-140): qubit by qubit, the
// probability of reading 0 given the bits chosen so far decides the next bit.  The reference re-parses and re-plans
// the full n-qubit circuit n times; here the circuit is planned ONCE (the plan does not depend on the measurement
// caps), compiled to one device plan, and the n evaluations only swap the rank-1 cap tensors.
// `contractionSequence`: a recorded plan in mCreatedFrom numbering (as produced by preProcess); empty -> in-process
// min-fill line-graph plan on the reduced circuit.
inline std::vector<bool> maxcutGetFinalString(const std::string &graphFilePath, int p, const std::vector<std::pair<int, int>> &contractionSequence,
                                              const std::vector<double> &gAndB, const std::string &outfilePath, unsigned tieSeed = 0,
                                              double *stringProbability = nullptr) {
    Timer clock;
    clock.start();
    ExtraData data(p, graphFilePath.c_str());
    const int n = data.numQubits;
    std::ostringstream circuit;
    circuit << n << std::endl;
    outputInitialPlusStateToFile(circuit, n);
    applyU_CsThenU_Bs(data.pairs, p, gAndB, n, circuit);
    std::string allTrace;
    for (int q = 0; q < n; ++q) allTrace += "T ";

    // plan on the host (no arithmetic), inputs snapshot in id order
    const bool before = device::Engine::PlanOnly();
    device::Engine::SetPlanOnly(true);
    std::vector<std::vector<std::complex<double>>> inputs;
    std::vector<qtb_plan_step> steps;
    long long units = 0;
    try {
        std::istringstream text(circuit.str());
        std::shared_ptr<Network> net = std::make_shared<Network>(text, allTrace);
        for (int i = 0; i < net->GetNumOriginalNodes(); ++i) inputs.push_back(net->GetAllNodes()[i]->GetTensorVals());
        if (contractionSequence.empty()) {
            net->ReduceCircuit();
            LineGraph lg(net);
            lg.SetQBBOutFiles("/dev/null", "", "/dev/null");          // ordering handed over in memory, no temp file
            lg.runMinFill();
            lg.LGContract();
        } else {
            ContractionTools tools(net);
            tools.ContractGivenSequence(contractionSequence);
        }
        units = net->getNumFloatOps();
        for (const auto &r : net->GetPlan()) {
            qtb_plan_step st;
            std::memset(&st, 0, sizeof(st));
            st.a = r.a; st.b = r.b; st.k = static_cast<int>(r.posA.size());
            for (int j = 0; j < st.k; ++j) { st.pos_a[j] = static_cast<int8_t>(r.posA[j]); st.pos_b[j] = static_cast<int8_t>(r.posB[j]); }
            steps.push_back(st);
        }
    } catch (...) {
        device::Engine::SetPlanOnly(before);
        throw;
    }
    device::Engine::SetPlanOnly(before);

    std::vector<int> ranks;
    for (const auto &in : inputs) { int r = 0; while ((static_cast<size_t>(1) << (2 * r)) < in.size()) ++r; ranks.push_back(r); }
    qtb_ctx *ctx = device::Engine::Get().ctx();
    qtb_plan *plan = nullptr;
    device::check(qtb_plan_create(ctx, static_cast<int>(ranks.size()), ranks.data(), static_cast<int>(steps.size()), steps.data(), &plan));

    const size_t firstCap = inputs.size() - static_cast<size_t>(n);        // the n measurement caps are the last original nodes
    const std::vector<std::complex<double>> capTrace = TraceNode().GetTensorVals(), capZero = ProjectZero().GetTensorVals(),
                                            capOne = ProjectOne().GetTensorVals();
    std::vector<bool> answer;
    std::mt19937 coin(tieSeed ? tieSeed : static_cast<unsigned>(std::time(nullptr)));
    double currentProb = 1.0;
    for (int q = 0; q < n; ++q) {
        for (int j = 0; j < n; ++j)
            inputs[firstCap + j] = j < static_cast<int>(answer.size()) ? (answer[j] ? capOne : capZero) : (j == q ? capZero : capTrace);
        std::vector<const double *> ptrs;
        for (const auto &in : inputs) ptrs.push_back(reinterpret_cast<const double *>(in.data()));
        double out[2] = {0.0, 0.0};
        device::check(qtb_plan_run_host(ctx, plan, ptrs.data(), out));
        const double probZero = out[0] / currentProb;                       // P(bit q = 0 | bits so far)
        if (probZero > 0.5) { currentProb *= probZero; answer.push_back(false); }
        else if (probZero < 0.5) { currentProb *= (1.0 - probZero); answer.push_back(true); }
        else { answer.push_back((coin() & 1u) != 0); currentProb *= 0.5; }
    }
    qtb_plan_destroy(ctx, plan);
    if (stringProbability) *stringProbability = currentProb;      // probability of the whole string (product of the conditionals)

    int cut = 0;
    for (const auto &e : data.pairs) if (answer[e.first] != answer[e.second]) ++cut;
    std::ofstream result(outfilePath);
    result << data.fileName << std::endl;
    for (bool b : answer) result << b << " ";
    result << std::endl << "Cut edges: " << cut << "/" << data.numQubits * 3 / 2 << std::endl;
    result << "Time elapsed: " << clock.getElapsed() << std::endl;
    if (!detail::quietMode()) std::cout << "Final string contracted with " << units << " units per qubit; cut edges: " << cut << std::endl;
    return answer;
}

}  // namespace qtorch
