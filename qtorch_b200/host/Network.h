// qtorch_b200/host/Network.h -- the tensor network of a circuit, with the contraction arithmetic on a B200.
//
// API surface of /root/reference/src/Network.h:54-140.  Host logic (parsing, ContractNodes bookkeeping,
// ReduceCircuit) reproduces the reference's sequence of decisions exactly, because the *contraction plan*
// -- node ids, wire order of every intermediate, the list of mCreatedFrom pairs -- must be identical to the
// reference's (SURVEY.md section 8a "plan-identity contract").  What changed:
//   * Network::ContractIndices (reference Network.h:876-971) no longer loops: it hands the step
//     (A, B, shared-leg positions) to libqtorch_b200 (qtb_contract), which runs it asynchronously on the GPU
//     -- tiny steps are grouped into one launch, large ones go to the DMMA tile kernel;
//   * the final value is read back lazily (one 16-byte D2H per network) instead of inside the hot loop;
//   * every executed step is recorded (GetPlan) so a network can be re-run as a compiled device plan.
#pragma once

#include <algorithm>
#include <cmath>
#include <csignal>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>

#include "DeviceEngine.h"
#include "Exceptions.h"
#include "Node.h"
#include "Timer.h"

namespace qtorch {

#define THRESH_RANK_THREAD 8   // rank from which the reference spawned threads; kept for its console message

// The reference exposes a process-wide watchdog (Network.h:50-51) that preProcess() arms.  Header-only,
// multi-TU-safe equivalents:
namespace detail {
template <class Dummy = void>
struct Globals {
    static Timer totTimer;
    static double maxTime;
};
template <class Dummy> Timer Globals<Dummy>::totTimer;
template <class Dummy> double Globals<Dummy>::maxTime = 60.0;

inline bool quietMode() {
    static const bool q = [] { const char *e = std::getenv("QTORCH_QUIET"); return e && std::atoi(e) != 0; }();
    return q;
}
}  // namespace detail
static Timer &totTimer = detail::Globals<>::totTimer;
static double &maxTime = detail::Globals<>::maxTime;

// One executed contraction step, in the reference's own numbering (ids = Node::mID).
struct PlanRecord {
    int a, b, c;                 // c = id given to the result (Network.h:853-857)
    int rankA, rankB, rankC;
    std::vector<int> posA, posB; // shared leg positions; posA increasing (Network.h:739-758)
};

class Network {
public:
    Network() {}
    Network(const std::string &inputFile, const std::string &measureFile) : mInputFile(inputFile), mMeasureFile(measureFile) {
        ParseNetwork(inputFile);
    }

    // addition: build from in-memory text (same grammar as the files): the term dispatcher hands circuits over
    // without the reference's write-file / re-parse round trip (maxcut.cpp:172-193)
    Network(std::istream &qasmText, const std::string &measurementText) : mInputFile("<memory>"), mMeasureFile("<memory>") {
        std::istringstream meas(measurementText);
        ParseStreams(qasmText, meas, true);
    }

    std::shared_ptr<Node> ContractNodes(std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB, int threshold);

    const std::complex<double> &GetFinalValue() const noexcept {
        const_cast<Network *>(this)->ResolveFinalValue();
        return mFinalVal;
    }
    void ContractNetworkLinearly();
    const std::vector<std::shared_ptr<Node>> &GetAllNodes() const noexcept { return mAllNodes; }
    void SetNumThreads(const int numThreads) noexcept { mNumberOfThreads = numThreads; }   // kept for API parity: the GPU ignores it
    const int GetNumQubits() const noexcept { return mNumberOfQubits; }
    const std::vector<std::shared_ptr<Node>> &GetUncontractedNodes() const noexcept {
        const_cast<Network *>(this)->CompactLive();
        return mUncontractedNodes;
    }
    const bool IsDone() noexcept { return mDone; }
    void MoveInitialStatesToBack();
    void ReduceCircuit();
    void OutputCircuitToVisualGraph(const std::string &toOutputTo) const;
    void OutputCircuitToTreewidthGraph(const std::string &toOutputTo) const;
    const bool HasFailed() const noexcept { return mFailure; }
    void Reset();
    const std::string &GetInputQasm() const noexcept { return mInputFile; }
    void resetFloatCounter() noexcept { mNumFloatOps = 0; }
    long long getNumFloatOps() noexcept { return mNumFloatOps; }

    // ---- additions (not in the reference) -------------------------------------------------------------
    const std::vector<PlanRecord> &GetPlan() const noexcept { return mPlan; }     // every executed step, in order
    int GetNumOriginalNodes() const noexcept { return mNumOriginalNodes; }          // nodes created by the parser
    static std::shared_ptr<Node> MakeMeasurementCap(char m, const char **what = nullptr);
    // start the device->host read of the final scalar behind the steps enqueued so far, without waiting for it; a later
    // GetFinalValue() then waits for that copy only.  Lets a caller enqueue the next network while this one still runs.
    void PrefetchFinalValue();
    ~Network() { DropPendingRead(); }

private:
    std::vector<std::shared_ptr<Node>> mNetworkParsingNodes;
    std::vector<std::shared_ptr<Wire>> mNetworkParsingWires;   // the dangling output wire of every qubit line while parsing
    std::string mInputFile, mMeasureFile;
    std::complex<double> mFinalVal{std::complex<double>(0.0)};
    std::shared_ptr<Node> mFinalSource;                        // rank-0 node whose scalar still has to be read back
    qtb_scalar_read *mFinalRead{nullptr};                      // ... or whose read-back is already in flight (PrefetchFinalValue)
    void DropPendingRead() noexcept {
        if (mFinalRead) { qtb_read_scalar_end(device::Engine::Get().ctx(), mFinalRead, nullptr); mFinalRead = nullptr; }
    }
    std::mutex mLocker;
    int mNumberOfQubits{0};
    int mDepth{0};
    bool mDone{false};
    bool mFailure{false};
    std::vector<std::shared_ptr<Node>> mAllNodes;
    std::vector<std::vector<std::shared_ptr<Node>>> mNodesByWire;
    // The reference erases operand A from this vector and puts C into B's slot on every step (Network.h:860-861), an
    // O(#nodes) shift per step.  Here the erase leaves a hole (nullptr) and the holes are squeezed out, order preserved,
    // the next time anybody LOOKS at the list (GetUncontractedNodes and the internal readers): planners that inspect
    // the list between steps see exactly the reference's vector, LGContract's 3000-step GHZ walk never pays the shifts.
    std::vector<std::shared_ptr<Node>> mUncontractedNodes;
    size_t mLiveHoles{0};
    size_t LiveCount() const noexcept { return mUncontractedNodes.size() - mLiveHoles; }
    void ReindexLive() noexcept {
        for (size_t i = 0; i < mUncontractedNodes.size(); ++i)
            if (mUncontractedNodes[i]) mUncontractedNodes[i]->mLiveSlot = static_cast<int>(i);
    }
    void CompactLive() noexcept {
        if (mLiveHoles == 0) return;
        size_t keep = 0;
        for (size_t i = 0; i < mUncontractedNodes.size(); ++i)
            if (mUncontractedNodes[i]) { if (keep != i) mUncontractedNodes[keep] = std::move(mUncontractedNodes[i]); ++keep; }
        mUncontractedNodes.resize(keep);
        mLiveHoles = 0;
        ReindexLive();
    }
    // operand A leaves the list, the result C takes operand B's place
    void RetireOperands(const std::shared_ptr<Node> &nodeA, const std::shared_ptr<Node> &nodeB, const std::shared_ptr<Node> &nodeC) {
        auto slotOf = [this](const std::shared_ptr<Node> &n) -> int {
            const int s = n->mLiveSlot;
            if (s >= 0 && s < static_cast<int>(mUncontractedNodes.size()) && mUncontractedNodes[s] == n) return s;
            CompactLive();                                   // not indexed (list edited from outside): search like the reference
            auto it = std::find(mUncontractedNodes.begin(), mUncontractedNodes.end(), n);
            return it == mUncontractedNodes.end() ? -1 : static_cast<int>(it - mUncontractedNodes.begin());
        };
        const int a = slotOf(nodeA);
        if (a >= 0) { mUncontractedNodes[a].reset(); ++mLiveHoles; nodeA->mLiveSlot = -1; }
        const int b = slotOf(nodeB);
        if (b >= 0) { mUncontractedNodes[b] = nodeC; nodeC->mLiveSlot = b; nodeB->mLiveSlot = -1; }
    }
    std::unordered_map<std::string, std::string> mArbitraryOneQubitGates, mArbitraryTwoQubitGates;
    long long mNumFloatOps{0};
    int mNumberOfThreads{8};
    std::vector<PlanRecord> mPlan;
    int mNumOriginalNodes{0};

    void ResolveFinalValue();
    void AttachOneQubitGate(const std::shared_ptr<Node> &gate, int qubit);
    void AttachTwoQubitGate(const std::shared_ptr<Node> &gate, int q1, int q2);
    int CheckedQubit(const std::string &token) const;

protected:
    inline void ContractIndices(const std::vector<std::pair<bool, int>> &toNotSumOn,
                                const std::vector<std::pair<int, int>> &toSumOn,
                                std::vector<int> &vectorIndexA, std::vector<int> &vectorIndexB,
                                std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB, std::shared_ptr<Node> nodeC);
    void ParseTokens(std::string &input, std::vector<std::string> &output);
    void ParseNetwork(const std::string &inputFile);
    void ParseStreams(std::istream &input, std::istream &measureStream, bool measureOpen);
    void ParseNode(std::string &inputLine);
    void CreateInitialStates();
    void AddMeasurementsOrTrace(std::vector<char> &measurements);
    void OutputCircuit(const std::vector<std::shared_ptr<Node>> &toOutput, const std::string &logFile) const;
    void FindAndReplace(std::vector<std::vector<std::shared_ptr<Node>>> &toSearch, std::shared_ptr<Node> toFind,
                        std::shared_ptr<Node> toReplaceWith) const;
    void FindAndReplace(std::vector<std::shared_ptr<Node>> &toSearch, std::shared_ptr<Node> toFind,
                        std::shared_ptr<Node> toReplaceWith) const;
    void FindAndRemove(std::vector<std::shared_ptr<Node>> &vect, std::shared_ptr<Node> toRemove) const;
};

// =====================================================================================================
// construction / parsing (reference Network.h:154-304, 326-707, 976-992)

inline void Network::Reset() {
    mNetworkParsingNodes.clear();
    mNetworkParsingWires.clear();
    mFinalVal = std::complex<double>(0.0);
    mFinalSource.reset();
    DropPendingRead();
    mNumberOfQubits = 0;
    mDepth = 0;
    mDone = false;
    mFailure = false;
    mAllNodes.clear();
    mNodesByWire.clear();
    mUncontractedNodes.clear();
    mLiveHoles = 0;
    mArbitraryOneQubitGates.clear();
    mArbitraryTwoQubitGates.clear();
    mPlan.clear();
    ParseNetwork(mInputFile);
}

inline void Network::CreateInitialStates() {
    for (int q = 0; q < mNumberOfQubits; ++q) {
        std::shared_ptr<Node> init = std::make_shared<ZeroStateNode>();
        if (!detail::quietMode()) std::cout << "Creating qubit " << q << " in the initial state: |0><0|" << std::endl;
        std::shared_ptr<Wire> out = std::make_shared<Wire>(init, nullptr, q);
        init->GetWires().push_back(out);
        mNetworkParsingWires.push_back(out);
        mNetworkParsingNodes.push_back(init);
        init->AddWireNumber(q);
        mNodesByWire[q].push_back(init);
        init->mID = static_cast<int>(mAllNodes.size());
        mAllNodes.push_back(init);
    }
}

// the rank-1 node that closes a qubit line for measurement character m (reference Network.h:199-237)
inline std::shared_ptr<Node> Network::MakeMeasurementCap(char m, const char **what) {
    const char *dummy = nullptr;
    const char *&w = what ? *what : dummy;
    switch (m) {
        case 'X': w = "Creating X measurement on qubit: "; return std::make_shared<XMeasure>();
        case 'Y': w = "Creating Y measurement on qubit: "; return std::make_shared<YMeasure>();
        case 'Z': w = "Creating Z measurement on qubit: "; return std::make_shared<ZMeasure>();
        case '0': w = "Creating Projection |0><0| measurement on qubit: "; return std::make_shared<ProjectZero>();
        case '1': w = "Creating Projection |1><1| measurement on qubit: "; return std::make_shared<ProjectOne>();
        default: w = "Tracing out qubit: "; return std::make_shared<TraceNode>();
    }
}

inline void Network::AddMeasurementsOrTrace(std::vector<char> &measurements) {
    for (int q = 0; q < mNumberOfQubits; ++q) {
        const char m = (static_cast<int>(measurements.size()) <= q) ? 'T' : measurements[q];
        const char *what = nullptr;
        std::shared_ptr<Node> cap = MakeMeasurementCap(m, &what);
        if (!detail::quietMode()) std::cout << what << q << std::endl;
        mNetworkParsingWires[q]->SetNodeB(cap);
        cap->GetWires().push_back(mNetworkParsingWires[q]);
        cap->AddWireNumber(q);
        mNodesByWire[q].push_back(cap);
        cap->mID = static_cast<int>(mAllNodes.size());
        mAllNodes.push_back(cap);
    }
}

inline void Network::ParseNetwork(const std::string &inputFile) {
    std::ifstream input(inputFile);
    if (!input.is_open()) {
        std::cout << "Failed to open QASM file!" << std::endl;
        mFailure = true;
        throw InvalidFile();
    }
    std::ifstream measureStream(mMeasureFile);
    ParseStreams(input, measureStream, measureStream.is_open());
}

inline void Network::ParseStreams(std::istream &input, std::istream &measureStream, bool measureOpen) {
    mAllNodes.reserve(4096);
    std::string line;
    std::getline(input, line);                 // first line: number of qubits
    mNumberOfQubits = std::stoi(line);
    mNodesByWire.resize(mNumberOfQubits);
    CreateInitialStates();

    if (!detail::quietMode()) std::cout << "Parsing nodes from file...." << std::endl;
    while (!input.eof()) {
        std::getline(input, line);
        ParseNode(line);
    }

    // measurement string: one character per qubit, whitespace ignored, missing entries trace the qubit out
    std::vector<char> measurements(mNumberOfQubits);
    if (!measureOpen) std::cout << "Measurement file failed to open - all qubits will be traced out" << std::endl;
    for (int q = 0; q < mNumberOfQubits; ++q) {
        char c;
        measurements[q] = (measureOpen && (measureStream >> c)) ? c : 'T';
    }
    AddMeasurementsOrTrace(measurements);

    mNetworkParsingNodes.clear();
    mNetworkParsingWires.clear();
    mUncontractedNodes = mAllNodes;
    mLiveHoles = 0;
    ReindexLive();
    mNumOriginalNodes = static_cast<int>(mAllNodes.size());
}

// split on single spaces; consecutive spaces yield empty tokens, a trailing space yields none
// (the behaviour of the reference's regex split, Network.h:988-991); whole-line '#' comments are skipped
inline void Network::ParseTokens(std::string &input, std::vector<std::string> &output) {
    if (input.empty() || input[0] == '#') return;
    size_t start = 0;
    while (true) {
        const size_t sp = input.find(' ', start);
        if (sp == std::string::npos) {
            if (start < input.size()) output.push_back(input.substr(start));
            break;
        }
        output.push_back(input.substr(start, sp - start));
        start = sp + 1;
    }
}

inline int Network::CheckedQubit(const std::string &token) const {
    const int q = std::stoi(token);
    if (q > mNumberOfQubits - 1 || q < 0) throw InvalidFileFormat();
    return q;
}

// gate on one line: consume the line's dangling wire, hang a fresh one on the output side
inline void Network::AttachOneQubitGate(const std::shared_ptr<Node> &gate, int qubit) {
    gate->GetWires().push_back(mNetworkParsingWires[qubit]);
    mNetworkParsingWires[qubit]->SetNodeB(gate);
    std::shared_ptr<Wire> out = std::make_shared<Wire>(gate, nullptr, qubit);
    mNetworkParsingWires[qubit] = out;
    gate->GetWires().push_back(out);
    gate->AddWireNumber(qubit);
    gate->mIndexOfPreviousNode = static_cast<int>(mNodesByWire[qubit].size()) - 1;
    mNodesByWire[qubit].push_back(gate);
}

// wire order of a two-qubit gate: [in_q1, in_q2, out_q1, out_q2] (reference Network.h:522-531)
inline void Network::AttachTwoQubitGate(const std::shared_ptr<Node> &gate, int q1, int q2) {
    const int qs[2] = {q1, q2};
    for (int q : qs) {
        gate->GetWires().push_back(mNetworkParsingWires[q]);
        mNetworkParsingWires[q]->SetNodeB(gate);
    }
    for (int q : qs) {
        std::shared_ptr<Wire> out = std::make_shared<Wire>(gate, nullptr, q);
        mNetworkParsingWires[q] = out;
        gate->GetWires().push_back(out);
    }
    for (int q : qs) {
        gate->AddWireNumber(q);
        mNodesByWire[q].push_back(gate);
    }
}

inline void Network::ParseNode(std::string &inputLine) {
    std::vector<std::string> tok;
    ParseTokens(inputLine, tok);
    if (tok.empty()) return;
    const std::string &op = tok[0];
    std::shared_ptr<Node> gate;

    auto twoQubits = [&](size_t first, int &q1, int &q2) {
        q1 = std::stoi(tok.at(first));
        q2 = std::stoi(tok.at(first + 1));
        if (q1 > mNumberOfQubits - 1 || q2 > mNumberOfQubits - 1 || q1 == q2 || q1 < 0 || q2 < 0) throw InvalidFileFormat();
    };

    // rotation angles go through float, as in the reference (std::stof, Network.h:342,383,402,422)
    if (op == "Rx" || op == "RX" || op == "Ry" || op == "RY" || op == "Rz" || op == "RZ" || op == "PHASE") {
        const float angle = std::stof(tok.at(1));
        if (op == "Rx" || op == "RX") gate = std::make_shared<RxNode>(angle);
        else if (op == "Ry" || op == "RY") gate = std::make_shared<RyNode>(angle);
        else if (op == "Rz" || op == "RZ") gate = std::make_shared<RzNode>(angle);
        else gate = std::make_shared<PhaseNode>(angle);
        AttachOneQubitGate(gate, CheckedQubit(tok.at(2)));
    } else if (op == "H" || op == "X" || op == "Y" || op == "Z") {
        if (op == "H") gate = std::make_shared<HNode>();
        else if (op == "X") gate = std::make_shared<XNode>();
        else if (op == "Y") gate = std::make_shared<YNode>();
        else gate = std::make_shared<ZNode>();
        AttachOneQubitGate(gate, CheckedQubit(tok.at(1)));
    } else if (op == "CNOT" || op == "SWAP" || op == "CRk" || op == "CZ") {
        int q1, q2;
        twoQubits(1, q1, q2);
        if (op == "CNOT") gate = std::make_shared<CNOTNode>();
        else if (op == "SWAP") gate = std::make_shared<SwapNode>();
        else if (op == "CRk") gate = std::make_shared<CRkNode>(q1);       // k := control qubit index (reference quirk)
        else gate = std::make_shared<CZNode>();
        AttachTwoQubitGate(gate, q1, q2);
    } else if (op == "CPHASE") {
        const double angle = std::stod(tok.at(1));                        // the one double-precision angle (Network.h:615)
        int q1, q2;
        twoQubits(2, q1, q2);
        gate = std::make_shared<CPhaseNode>(angle);
        AttachTwoQubitGate(gate, q1, q2);
    } else if (op == "def1") {
        mArbitraryOneQubitGates.insert({tok.at(1), tok.at(2)});
        return;
    } else if (op == "def2") {
        mArbitraryTwoQubitGates.insert({tok.at(1), tok.at(2)});
        return;
    } else if (mArbitraryOneQubitGates.count(op)) {
        gate = std::make_shared<ArbitraryOneQubitNode>(mArbitraryOneQubitGates[op], op);
        AttachOneQubitGate(gate, CheckedQubit(tok.at(1)));
    } else if (mArbitraryTwoQubitGates.count(op)) {
        gate = std::make_shared<ArbitraryTwoQubitNode>(mArbitraryTwoQubitGates[op], op);
        int q1, q2;
        twoQubits(1, q1, q2);
        AttachTwoQubitGate(gate, q1, q2);
    } else {
        std::cout << "Failed to compile line: " << std::endl;
        for (const auto &t : tok) std::cout << t << " ";
        std::cout << std::endl;
        throw InvalidFileFormat();
    }
    gate->mID = static_cast<int>(mAllNodes.size());
    mAllNodes.push_back(gate);
}

// =====================================================================================================
// the step: bookkeeping on the host, arithmetic on the device

inline std::shared_ptr<Node> Network::ContractNodes(std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB, int threshold) {
    std::unique_lock<std::mutex> guard(mLocker);
    if (nodeA->mContracted || nodeB->mContracted) return nullptr;

    auto touches = [](const std::shared_ptr<Wire> &w, const std::shared_ptr<Node> &n) {
        return w->GetNodeB().lock() == n || w->GetNodeA().lock() == n;
    };

    // legs of A: shared with B (summed, kept in A-wire order) or free (become C's first legs)
    std::vector<int> sharedA, sharedB;
    std::vector<std::pair<bool, int>> freeLegs;                 // (true, i): leg i of A; (false, j): leg j of B
    std::vector<std::shared_ptr<Wire>> sharedWires, keptWires;
    const std::vector<std::shared_ptr<Wire>> &wiresA = nodeA->GetWires();
    const std::vector<std::shared_ptr<Wire>> &wiresB = nodeB->GetWires();
    for (int i = 0; i < static_cast<int>(wiresA.size()); ++i) {
        if (touches(wiresA[i], nodeB)) { sharedA.push_back(i); sharedWires.push_back(wiresA[i]); }
        else { freeLegs.push_back({true, i}); keptWires.push_back(wiresA[i]); }
    }
    for (const auto &w : sharedWires)
        for (int j = 0; j < static_cast<int>(wiresB.size()); ++j)
            if (w == wiresB[j]) sharedB.push_back(j);
    for (int j = 0; j < static_cast<int>(wiresB.size()); ++j) {
        if (!touches(wiresB[j], nodeA)) { freeLegs.push_back({false, j}); keptWires.push_back(wiresB[j]); }
    }

    // rejection rule (reference Network.h:772-777): unconnected (unless both are scalars) or result too large
    const bool unconnected = sharedA.empty() && !freeLegs.empty();
    if (unconnected || static_cast<int>(keptWires.size()) > std::max(nodeA->mRank, nodeB->mRank) + threshold) return nullptr;

    std::vector<std::pair<int, int>> sharedPairs(sharedA.size());
    for (size_t j = 0; j < sharedA.size(); ++j) sharedPairs[j] = {sharedA[j], sharedB[j]};

    // node C takes over the kept wires, A's first (reference Network.h:809-818)
    std::shared_ptr<Node> nodeC = std::make_shared<Node>(static_cast<int>(freeLegs.size()));
    for (const auto &w : keptWires) {
        nodeC->GetWires().push_back(w);
        const std::shared_ptr<Node> endA = w->GetNodeA().lock();
        if (endA == nodeA || endA == nodeB) w->SetNodeA(nodeC);
        else {
            const std::shared_ptr<Node> endB = w->GetNodeB().lock();
            if (endB == nodeA || endB == nodeB) w->SetNodeB(nodeC);
        }
    }
    std::vector<int> digitsA(nodeA->mRank), digitsB(nodeB->mRank);
    if (nodeC->mRank >= THRESH_RANK_THREAD && !detail::quietMode())
        std::cout << "Contracting Nodes of Rank " << nodeA->mRank << " and " << nodeB->mRank
                  << " to get a Node of Rank: " << nodeC->mRank << " Hold On....." << std::endl;

    nodeB->mContracted = true;
    nodeA->mContracted = true;
    for (const auto &w : sharedWires) w->SetIsContracted(true);
    guard.unlock();

    ContractIndices(freeLegs, sharedPairs, digitsA, digitsB, nodeA, nodeB, nodeC);   // enqueue on the GPU

    PlanRecord rec;
    rec.a = nodeA->mID; rec.b = nodeB->mID;
    rec.rankA = nodeA->mRank; rec.rankB = nodeB->mRank; rec.rankC = nodeC->mRank;
    rec.posA = sharedA; rec.posB = sharedB;

    if (mDone) {
        // last step of the network: the reference files a rank-0 bookkeeping node and returns nullptr
        std::shared_ptr<Node> marker = std::make_shared<Node>(0);
        marker->mID = static_cast<int>(mAllNodes.size());
        marker->mCreatedFrom = {nodeA->mID, nodeB->mID};
        rec.c = marker->mID;
        mPlan.push_back(rec);
        mAllNodes.push_back(marker);
        RetireOperands(nodeA, nodeB, nodeC);
        nodeA->ClearNodeData();
        nodeB->ClearNodeData();
        return nullptr;
    }
    guard.lock();
    nodeC->mID = static_cast<int>(mAllNodes.size());
    nodeC->mCreatedFrom = {nodeA->mID, nodeB->mID};
    rec.c = nodeC->mID;
    mPlan.push_back(rec);
    mAllNodes.push_back(nodeC);
    nodeA->ClearNodeData();           // stream-ordered free: the step that reads them is already enqueued
    nodeB->ClearNodeData();
    RetireOperands(nodeA, nodeB, nodeC);                  // C inherits B's slot
    return nodeC;
}

// Replaces the reference's index-arithmetic double loop (Network.h:892-960) by one asynchronous device step.
inline void Network::ContractIndices(const std::vector<std::pair<bool, int>> &toNotSumOn,
                                     const std::vector<std::pair<int, int>> &toSumOn,
                                     std::vector<int> &vectorIndexA, std::vector<int> &vectorIndexB,
                                     std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB, std::shared_ptr<Node> nodeC) {
    (void)vectorIndexA; (void)vectorIndexB;
    // the reference's "float op" counter: 4^(free + summed legs) per step (Network.h:884-885)
    const int independent = static_cast<int>(toNotSumOn.size() + toSumOn.size());
    mNumFloatOps += static_cast<long long>(std::pow(4, independent));

    if (!nodeA->HasData() || !nodeB->HasData()) throw InvalidFunctionInput();     // Network.h:938-940

    if (!device::Engine::PlanOnly()) {
        int posA[QTB_MAX_RANK], posB[QTB_MAX_RANK];
        const int k = static_cast<int>(toSumOn.size());
        if (k > QTB_MAX_RANK || nodeC->mRank > QTB_MAX_RANK) throw ContractionFailure();
        for (int j = 0; j < k; ++j) { posA[j] = toSumOn[j].first; posB[j] = toSumOn[j].second; }
        // the watchdog of the reference interrupts a step half-way (Network.h:899); device steps are atomic,
        // so it is only consulted between steps
        if (totTimer.getElapsed() < maxTime) {
            const qtb_tensor ta = nodeA->DeviceTensor(), tb = nodeB->DeviceTensor(), tc = nodeC->DeviceOutput();
            device::check(qtb_contract(device::Engine::Get().ctx(), ta, tb, k, posA, posB, tc));
        }
    }

    if (toNotSumOn.empty()) {
        // rank-0 result.  Reference rule (Network.h:961-969): it becomes the final value if no non-zero
        // final value exists yet, or if this is a product of two scalars (disconnected components).
        ResolveFinalValue();
        if ((std::abs(mFinalVal.real()) <= 1.0e-30 && std::abs(mFinalVal.imag()) <= 1.0e-30) ||
            (nodeA->mRank == 0 && nodeB->mRank == 0))
            mFinalSource = nodeC;                      // read back lazily: no sync inside the contraction
        if (LiveCount() == 2) mDone = true;
    }
}

inline void Network::PrefetchFinalValue() {
    if (!mFinalSource || mFinalRead || device::Engine::PlanOnly() || !mFinalSource->OnDevice()) return;
    device::check(qtb_read_scalar_begin(device::Engine::Get().ctx(), mFinalSource->DeviceTensor(), &mFinalRead));
    mFinalSource.reset();
}

inline void Network::ResolveFinalValue() {
    if (mFinalRead) {
        double v[2] = {0.0, 0.0};
        qtb_scalar_read *r = mFinalRead;
        mFinalRead = nullptr;
        device::check(qtb_read_scalar_end(device::Engine::Get().ctx(), r, v));
        mFinalVal = std::complex<double>(v[0], v[1]);
    }
    if (!mFinalSource) return;
    std::shared_ptr<Node> src;
    src.swap(mFinalSource);
    if (device::Engine::PlanOnly()) { mFinalVal = std::complex<double>(std::numeric_limits<double>::quiet_NaN(), 0.0); return; }
    if (src->OnDevice()) {
        double v[2] = {0.0, 0.0};
        device::check(qtb_read_scalar(device::Engine::Get().ctx(), src->DeviceTensor(), v));
        mFinalVal = std::complex<double>(v[0], v[1]);
    } else if (src->HasData()) {
        mFinalVal = src->Access(0);
    }
}

inline void Network::ContractNetworkLinearly() {
    while (!mDone) {
        CompactLive();
        const std::shared_ptr<Wire> &w = mUncontractedNodes[0]->GetWires()[0];
        if (w->GetNodeA().expired() || w->GetNodeB().expired()) throw InvalidContractionMethod();
        ContractNodes(w->GetNodeA().lock(), w->GetNodeB().lock(), 1000);
    }
}

inline void Network::MoveInitialStatesToBack() {
    const int n = mNumberOfQubits;
    auto rotate = [n](std::vector<std::shared_ptr<Node>> &v) {
        const long last = static_cast<long>(v.size()) - 1 - n, first = static_cast<long>(v.size()) - 1 - 2L * n;
        int count = 0;
        for (long i = last; i >= first && i >= 0; --i) std::swap(v[count++], v[static_cast<size_t>(i)]);
    };
    rotate(mAllNodes);
    CompactLive();
    rotate(mUncontractedNodes);
    ReindexLive();
}

// =====================================================================================================
// ReduceCircuit (reference Network.h:1015-1118): same sequence of ContractNodes calls, hence same plan.
//  pass 1: every one-qubit gate is absorbed into the node before it on its line (threshold 0);
//  pass 2: walk all lines in lock-step and fuse consecutive two-qubit gates acting on the same qubit pair.

inline void Network::ReduceCircuit() {
    std::vector<std::vector<std::shared_ptr<Node>>> kept(mNumberOfQubits);
    for (int q = 0; q < mNumberOfQubits; ++q) {
        std::vector<std::shared_ptr<Node>> &line = mNodesByWire[q];
        for (size_t pos = 0; pos < line.size(); ++pos) {
            std::shared_ptr<Node> gate = line[pos];
            if (gate->mRank != 2) { kept[q].push_back(gate); continue; }
            std::shared_ptr<Node> before = kept[q].back();
            std::shared_ptr<Node> fused = ContractNodes(before, gate, 0);
            // the fused node inherits the qubit labels of the node it grew from (the gate's input wire still
            // names that node as its A end)
            const std::shared_ptr<Node> origin = gate->GetWires()[0]->GetNodeA().lock();
            fused->AddWireNumber(origin->GetWireNumber()[0]);
            if (origin->mRank > 2) fused->AddWireNumber(origin->GetWireNumber()[1]);
            if (before->GetTypeOfNode() == GateType::INITSTATE) {
                fused->SetTypeOfNode(GateType::INITSTATE);
                fused->SetTypeOfNodeString("INITSTATE(Manipulated)");
            }
            // `before` only ever sits on its own qubit lines, so searching those is equivalent to the reference's
            // scan over every line (Network.h:1034-1035) at O(1) instead of O(#qubits) per merge
            for (int line : before->GetWireNumber()) {
                if (line < 0 || line >= mNumberOfQubits) continue;
                FindAndReplace(mNodesByWire[line], before, fused);
                FindAndReplace(kept[line], before, fused);
            }
        }
    }
    mNodesByWire = std::move(kept);

    const int nq = mNumberOfQubits;
    std::vector<std::vector<std::shared_ptr<Node>>> merged(nq);
    std::vector<std::shared_ptr<Node>> head(nq);      // last node placed on each line
    std::vector<int> cursor(nq, 0);
    std::vector<char> advanced(nq, 0);
    bool fusedThisSweep = false;
    // The reference sweeps over every line again and again, each line advancing by at most one node per sweep
    // (Network.h:1050-1112).  A line that could not advance stays blocked until the line it waits for moves, so a sweep
    // only has to visit (in ascending order, like the reference) the lines that advanced in the previous sweep and the
    // lines of the nodes those now have in front of them: same decisions in the same order, O(#nodes log) instead of
    // O(depth x qubits) -- the 1000-qubit GHZ chain has depth 1000.
    std::vector<int> visit(nq), movedLines, next;
    std::vector<char> queued(nq, 0);
    for (int q = 0; q < nq; ++q) visit[q] = q;
    auto pending = [&](int q) { return cursor[q] < static_cast<int>(mNodesByWire[q].size()); };
    while (!visit.empty()) {
        movedLines.clear();
        for (int q : visit) {
            if (!pending(q) || advanced[q]) continue;
            Node *peek = mNodesByWire[q][cursor[q]].get();       // no shared_ptr copy in this (hot) scan
            if (peek->mRank == 1) {
                head[q] = mNodesByWire[q][cursor[q]];
                advanced[q] = 1;
                movedLines.push_back(q);
                ++cursor[q];
                continue;
            }
            const int qa = peek->GetWireNumber()[0], qb = peek->GetWireNumber()[1];
            if (advanced[qb] || advanced[qa]) continue;
            // the gate must be next on BOTH of its lines
            if (mNodesByWire[qa][cursor[qa]].get() != mNodesByWire[qb][cursor[qb]].get()) continue;
            std::shared_ptr<Node> cand = mNodesByWire[q][cursor[q]];
            if (head[qa] == head[qb]) {
                // previous node on both lines is one and the same two-qubit node: fuse (later gate is operand A)
                head[qa] = ContractNodes(cand, head[qa], 0);
                head[qb] = head[qa];
                head[qa]->AddWireNumber(qa);
                head[qa]->AddWireNumber(qb);
                fusedThisSweep = true;
            } else {
                head[qa] = cand;
                head[qb] = cand;
            }
            ++cursor[qa]; ++cursor[qb];
            advanced[qa] = advanced[qb] = 1;
            movedLines.push_back(qa);
            movedLines.push_back(qb);
        }
        // The reference pads waiting lines with a nullptr per sweep (Network.h:1102-1104).  Every reader of
        // mNodesByWire skips nullptr entries (Network.h:1179,1217), so the padding is unobservable; it is left
        // out because it costs O(depth x qubits) memory and time (16 MB for the 1000-qubit GHZ chain).
        for (int q : movedLines) {
            if (!advanced[q]) continue;                           // (a line can be listed twice)
            if (fusedThisSweep && !merged[q].empty()) merged[q].back() = head[q];
            else merged[q].push_back(head[q]);
            advanced[q] = 0;
        }
        fusedThisSweep = false;
        // who can possibly advance in the next sweep: the lines that moved, and both lines of the node each now faces
        next.clear();
        auto enqueue = [&](int q) { if (q >= 0 && q < nq && !queued[q] && pending(q)) { queued[q] = 1; next.push_back(q); } };
        for (int q : movedLines) {
            enqueue(q);
            if (pending(q)) {
                Node *front = mNodesByWire[q][cursor[q]].get();
                if (front->mRank != 1) { enqueue(front->GetWireNumber()[0]); enqueue(front->GetWireNumber()[1]); }
            }
        }
        std::sort(next.begin(), next.end());
        for (int q : next) queued[q] = 0;
        visit.swap(next);
    }
    mNodesByWire = std::move(merged);
}

// =====================================================================================================
// graph dumps and small helpers (reference Network.h:1122-1282)

inline void Network::OutputCircuit(const std::vector<std::shared_ptr<Node>> &toOutput, const std::string &logFile) const {
    std::ofstream out(logFile);
    if (!out.is_open()) {
        std::cout << "Failure to Output Circuit to File" << std::endl;
        throw InvalidFile();
    }
    for (const auto &n : toOutput) {
        out << n->GetTypeOfNodeString() << " ";
        for (int q : n->GetWireNumber()) out << q << " ";
        out << std::endl;
    }
    if (mDepth != 0) out << "Depth: " << mDepth << std::endl;
    else out << "Depth Has Not Been Calculated due to Non-Local Interactions" << std::endl;
}

inline void Network::OutputCircuitToVisualGraph(const std::string &toOutputTo) const {
    std::ofstream out(toOutputTo);
    if (!out.is_open()) {
        std::cout << "Failure to Output Circuit to Visual Graph" << std::endl;
        throw InvalidFile();
    }
    std::unordered_map<std::shared_ptr<Node>, int> number;
    out << "graph " << mInputFile.substr(0, mInputFile.find('.')) << "{" << std::endl;
    out << "node [height=1, width=.1];\n rankdir=LR;" << std::endl;
    int next = 0;
    for (const auto &n : GetUncontractedNodes()) {
        out << "node" << next << " [label=\"" << n->GetTypeOfNodeString() << "\"";
        if (n->mRank == 1) out << ", height = .5";
        out << "];" << std::endl;
        number.insert({n, next++});
    }
    for (const auto &line : mNodesByWire) {
        size_t prev = 0;
        for (size_t i = 1; i < line.size(); ++i) {
            if (line[i] == nullptr) continue;
            out << "node" << number[line[prev]] << " -- node" << number[line[i]] << std::endl;
            prev = i;
        }
    }
    out << "}" << std::endl;
}

inline void Network::OutputCircuitToTreewidthGraph(const std::string &toOutputTo) const {
    std::ofstream out(toOutputTo);
    if (!out.is_open()) {
        std::cout << "Failure to Output Circuit to TW Graph" << std::endl;
        throw InvalidFile();
    }
    out << "c Created From File: " << mInputFile << std::endl;
    std::unordered_map<std::shared_ptr<Node>, int> number;
    int next = 0;
    for (const auto &n : GetUncontractedNodes()) number.insert({n, next++});
    for (size_t l = 0; l < mNodesByWire.size(); ++l) {
        const auto &line = mNodesByWire[l];
        size_t prev = 0;
        for (size_t i = 1; i < line.size(); ++i) {
            if (line[i] == nullptr) continue;
            out << "e " << number[line[prev]] << " " << number[line[i]];
            const bool veryLast = (i == line.size() - 1 && l == mNodesByWire.size() - 1);
            if (!veryLast) out << std::endl;
            prev = i;
        }
    }
}

inline void Network::FindAndReplace(std::vector<std::vector<std::shared_ptr<Node>>> &toSearch, std::shared_ptr<Node> toFind,
                                    std::shared_ptr<Node> toReplaceWith) const {
    int hits = 0;
    for (auto &line : toSearch) {
        for (auto &slot : line) {
            if (slot == toFind) { slot = toReplaceWith; ++hits; }
            if (hits >= 2) break;
        }
    }
}

inline void Network::FindAndReplace(std::vector<std::shared_ptr<Node>> &toSearch, std::shared_ptr<Node> toFind,
                                    std::shared_ptr<Node> toReplaceWith) const {
    auto it = std::find(toSearch.begin(), toSearch.end(), toFind);
    if (it != toSearch.end()) *it = toReplaceWith;
}

inline void Network::FindAndRemove(std::vector<std::shared_ptr<Node>> &vect, std::shared_ptr<Node> toRemove) const {
    auto it = std::find(vect.begin(), vect.end(), toRemove);
    if (it != vect.end()) vect.erase(it);
}

}  // namespace qtorch
