// qtorch_b200/host/ContractionTools.h -- host-side planners that end in Network::ContractNodes calls.
// Interface of /root/reference/src/ContractionTools.h:52-119.  The planners are host logic and keep the
// reference's decision rules (thresholds, escalation, partitioning); only the arithmetic behind
// ContractNodes moved to the GPU.  Implemented: Stochastic (ParallelContract, reference :219-388),
// FromEdges (:1072-1190), ContractGivenSequence (:392-411, the plan-replay entry),
// ContractUserDefinedSequenceOfWires (:154-214), the reduce-and-print helpers (:1196-1229) and
// CalculateTreewidth (:1232-1256) and the sampled greedy search CostContractSimple (:837-1046: draw ~log2(n) connected
// pairs, score each by the 4^(rA+rB-k) cost of the step plus pValue randomly grown look-ahead steps, contract the
// cheapest; rank / wire thresholds escalate when every draw is rejected).  With the same seed it draws the same random
// numbers in the same order as the reference, so the plans coincide (tests/golden "cost" cases).  The multi-threaded
// brute-force sampler (:431-835) is not on the hot path and is not provided: it throws InvalidContractionMethod.
//
// Addition: SetSeed() / QTORCH_SEED make the stochastic search reproducible (the reference seeds from
// std::random_device, ContractionTools.h:61, so its plans cannot be replayed without recording them).
#pragma once

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <thread>

#include "Exceptions.h"
#include "LineGraph.h"
#include "Network.h"

namespace qtorch {

enum ContractionType { Stochastic, FromEdges, CostContractSimple, CostContractBruteForce };

class ContractionTools {
public:
    explicit ContractionTools(const std::string &inputFile, const std::string &measureFile, const int numThreads = 8)
        : mString(inputFile), mMeasureFile(measureFile), mCopyCreated(false), mNumThreadsInNetwork(numThreads) { SeedGenerator(); }
    explicit ContractionTools(const std::shared_ptr<Network> network) : mCopyCreated(true) {
        SeedGenerator();
        mNetwork = network;
    }

    std::complex<double> GetFinalVal() noexcept { return mFinalVal; }
    std::shared_ptr<Network> Contract(ContractionType type, int pValue = 1, int numSamples = 1);
    std::shared_ptr<Network> ReduceAndPrintCircuitToVisualGraph(const std::string &toPrintTo) const;
    std::shared_ptr<Network> ReduceAndPrintCircuitToTWGraph(const std::string &toPrintTo) const;
    std::shared_ptr<Network> ContractUserDefinedSequenceOfWires(const std::string &userInputFilePath);
    const int CalculateTreewidth(const int qbbseconds, const bool sixtyFourBitOpSystem = true) const;
    std::shared_ptr<Network> ContractGivenSequence(const std::vector<std::pair<int, int>> &sequence);

    void Reset(const std::string &inputFile, const std::string &measureFile, const int numThreads = 8) {
        mNetwork = nullptr; mCopyCreated = false;
        mString = inputFile; mMeasureFile = measureFile; mNumThreadsInNetwork = numThreads;
    }
    void Reset() { mNetwork = nullptr; mCopyCreated = false; }
    void Reset(std::shared_ptr<Network> network) { mNetwork = network; mCopyCreated = true; }

    void SetSeed(unsigned seed) { mRandGen = std::mt19937(seed); }     // addition

private:
    std::string mString, mMeasureFile;
    std::complex<double> mFinalVal;
    std::vector<std::shared_ptr<std::vector<std::shared_ptr<Node>>>> mPartitionedNodes;
    std::random_device mRandDevice;
    std::shared_ptr<Network> mNetwork;
    std::mt19937 mRandGen;
    bool mCopyCreated;
    int mNumThreadsInNetwork{8};

    void SeedGenerator() {
        if (const char *e = std::getenv("QTORCH_SEED")) mRandGen = std::mt19937(static_cast<unsigned>(std::strtoul(e, nullptr, 10)));
        else mRandGen = std::mt19937(mRandDevice());
    }
    std::shared_ptr<Network> OpenNetwork() const {
        if (mCopyCreated) return mNetwork;
        std::shared_ptr<Network> net = std::make_shared<Network>(mString, mMeasureFile);
        net->SetNumThreads(mNumThreadsInNetwork);
        return net;
    }

protected:
    void CreateChunksOfNodes(std::shared_ptr<Network> &myNetwork);
    std::shared_ptr<Network> ParallelContract(std::mt19937 &randomGenerator);
    std::shared_ptr<Network> ContractFromEdges(std::mt19937 &randomGenerator);
    std::shared_ptr<Network> CostBasedContractionSimple(const int pValue);
    long long CalculateCost(const int pVal, const int indexA, const int indexB, const int thresholdFinalRank, const int thresholdNumwires);
    int NumberOfConnectedBetweenTwoSuperNodes(const std::vector<std::shared_ptr<Node>> &groupA, const std::vector<std::shared_ptr<Node>> &groupB) {
        int n = 0;
        for (const auto &a : groupA)
            for (const auto &b : groupB) n += NumberOfConnectedWires(a, b);
        return n;
    }
    int NumberOfConnectedWires(std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB) {
        int n = 0;
        for (const auto &w : nodeA->GetWires())
            if (w->GetNodeA().lock() == nodeB || w->GetNodeB().lock() == nodeB) ++n;
        return n;
    }
};

inline std::shared_ptr<Network> ContractionTools::Contract(ContractionType type, int pValue, int numSamples) {
    (void)numSamples;
    if (type == Stochastic) return ParallelContract(mRandGen);
    if (type == FromEdges) return ContractFromEdges(mRandGen);
    if (type == CostContractSimple) return CostBasedContractionSimple(pValue);
    throw InvalidContractionMethod();        // brute-force sampler: not provided (see header comment)
}

// ---- CostContractSimple (reference ContractionTools.h:837-899) ----------------------------------------
// One step per round: about log2(#nodes) admissible random pairs are scored and the cheapest is contracted.
// A draw is thrown away (and redrawn) when the two nodes are not connected, are the same node, or are the
// pair currently held as best; a draw whose score is -1 (over the rank / wire thresholds) counts as a
// failure, and more failures than nodes relax both thresholds by one.  The order of the random draws is
// part of the plan identity, so the control flow below follows the reference statement by statement.
inline std::shared_ptr<Network> ContractionTools::CostBasedContractionSimple(const int pValue) {
    if (!mCopyCreated) mNetwork = OpenNetwork();
    if (mNetwork->HasFailed()) return nullptr;
    while (!mNetwork->IsDone() && totTimer.getElapsed() < maxTime) {
        const auto &live = mNetwork->GetUncontractedNodes();
        long long best = -1;
        int pick[2] = {0, 1};
        std::uniform_int_distribution<> draw(0, static_cast<int>(live.size()) - 1);
        int rejected = 0, rankLimit = 11, wireLimit = 8;
        if (live.size() != 2) {
            for (int round = 0; round < std::log2(live.size()) && totTimer.getElapsed() < maxTime; round++) {
                const int one = draw(mRandGen);
                const int two = draw(mRandGen);
                const bool heldPair = (pick[0] == one && pick[1] == two) || (pick[1] == one && pick[0] == two);
                if (NumberOfConnectedWires(live[one], live[two]) == 0 || heldPair || one == two) { --round; continue; }
                const long long score = CalculateCost(pValue, one, two, rankLimit, wireLimit);
                if (score == -1) {
                    ++rejected;
                    --round;
                    if (rejected > static_cast<int>(live.size())) { ++rankLimit; ++wireLimit; rejected = 0; }
                    continue;
                }
                if (score < best || round == 0) { pick[0] = one; pick[1] = two; best = score; }
            }
        }
        mNetwork->ContractNodes(live[pick[0]], live[pick[1]], 1000000);
        if (!detail::quietMode()) std::cout << "Nodes Left: " << mNetwork->GetUncontractedNodes().size() << std::endl;
    }
    if (!mNetwork->IsDone()) throw ContractionFailure();
    mFinalVal = mNetwork->GetFinalValue();
    return mNetwork;
}

// Score of contracting live nodes indexA, indexB (reference ContractionTools.h:901-1046).
//   pVal == 0: rank(A) + rank(B) - (wires between them).
//   otherwise: -1 if the result rank exceeds thresholdFinalRank or more than thresholdNumwires wires connect the pair;
//   else 4^(rA+rB-k) plus the cost of pVal further steps, each absorbing one randomly drawn neighbour of the growing
//   super-node (a neighbour that would push the work exponent to >= 12 is redrawn up to 2*|neighbours| times).
// The "selected" marks live on the nodes (Node::mSelectedInCostContractionAlgorithm) and the reference leaves them set
// on its two early exits (no neighbours left / too many redraws); later calls see those marks, so the exits are kept.
inline long long ContractionTools::CalculateCost(const int pVal, const int indexA, const int indexB, const int thresholdFinalRank,
                                                 const int thresholdNumwires) {
    const auto &live = mNetwork->GetUncontractedNodes();
    const std::shared_ptr<Node> &a = live[indexA], &b = live[indexB];
    int shared = NumberOfConnectedWires(a, b);
    if (pVal == 0) return a->mRank + b->mRank - shared;
    if (a->mRank + b->mRank - 2 * shared > thresholdFinalRank || shared > thresholdNumwires) return -1;

    std::vector<std::shared_ptr<Node>> super, frontier;
    // every unmarked node on the far side of one of n's wires joins the frontier (and gets marked)
    auto growFrontier = [&frontier](const std::shared_ptr<Node> n) {       // by value: frontier may reallocate while n's wires are walked
        for (const auto &w : n->GetWires()) {
            const std::shared_ptr<Node> endA = w->GetNodeA().lock(), endB = w->GetNodeB().lock();
            if (endA->mSelectedInCostContractionAlgorithm && endB->mSelectedInCostContractionAlgorithm) continue;
            const std::shared_ptr<Node> &fresh = endA->mSelectedInCostContractionAlgorithm ? endB : endA;
            fresh->mSelectedInCostContractionAlgorithm = true;
            frontier.push_back(fresh);
        }
    };
    long long cost = static_cast<long long>(std::pow(4, a->mRank + b->mRank - shared));
    int superRank = a->mRank + b->mRank - 2 * shared;
    super.push_back(a); a->mSelectedInCostContractionAlgorithm = true;
    super.push_back(b); b->mSelectedInCostContractionAlgorithm = true;
    growFrontier(a);
    growFrontier(b);
    int redraws = 0;
    for (int step = 0; step < pVal; step++) {
        if (frontier.empty()) return cost;                                     // (marks stay set, as in the reference)
        std::uniform_int_distribution<> draw(0, static_cast<int>(frontier.size()) - 1);
        // slots of absorbed neighbours are empty.  Deviation: when EVERY slot is empty (the super-node has swallowed its
        // whole neighbourhood, typical for pValue >= 2) the reference redraws forever (:1002-1004); stop looking ahead.
        if (std::all_of(frontier.begin(), frontier.end(), [](const std::shared_ptr<Node> &n) { return n == nullptr; })) break;
        int r = draw(mRandGen);
        while (frontier[r] == nullptr) r = draw(mRandGen);
        shared = NumberOfConnectedBetweenTwoSuperNodes(super, {frontier[r]});
        if (superRank + frontier[r]->mRank - shared >= 12 && redraws < static_cast<int>(frontier.size()) * 2) {
            ++redraws;
            --step;
            continue;
        }
        if (redraws > static_cast<int>(frontier.size()) * 2) return -1;        // (marks stay set, as in the reference)
        redraws = 0;
        cost += std::pow(4, superRank) * std::pow(4, frontier[r]->mRank) / std::pow(4, shared);
        superRank += frontier[r]->mRank - 2 * shared;
        growFrontier(frontier[r]);
        super.push_back(std::move(frontier[r]));
    }
    for (auto &n : super) n->mSelectedInCostContractionAlgorithm = false;
    for (auto &n : frontier)
        if (n != nullptr) n->mSelectedInCostContractionAlgorithm = false;
    return cost;
}

// remove the two operands (positions one, two) from a working list the way the reference does -- the
// order of the swap-with-back moves decides which node ends up where, and with it the later random picks
namespace detail {
inline void dropPair(std::vector<std::shared_ptr<Node>> &v, int one, int two) {
    if (two == static_cast<int>(v.size()) - 1) {
        v[two] = v.back(); v.pop_back();
        v[one] = v.back(); v.pop_back();
    } else {
        v[one] = v.back(); v.pop_back();
        v[two] = v.back(); v.pop_back();
    }
}
}  // namespace detail

inline void ContractionTools::CreateChunksOfNodes(std::shared_ptr<Network> &myNetwork) {
    // two partitions of equal size (the remainder forms further chunks); note the list is NOT cleared between
    // calls, exactly like the reference (ContractionTools.h:1261-1279)
    const std::vector<std::shared_ptr<Node>> &nodes = myNetwork->GetUncontractedNodes();
    const int numPartitions = 2;
    const int size = static_cast<int>(nodes.size());
    int take = size / numPartitions, done = 0;
    size_t at = 0;
    while (at < nodes.size()) {
        auto part = std::make_shared<std::vector<std::shared_ptr<Node>>>(nodes.begin() + at, nodes.begin() + at + take);
        mPartitionedNodes.push_back(part);
        at += take;
        done += take;
        if (size - done < take) take = size - done;
        if (take <= 0) break;
    }
}

inline std::shared_ptr<Network> ContractionTools::ParallelContract(std::mt19937 &randomGenerator) {
    std::shared_ptr<Network> net = OpenNetwork();
    if (net->HasFailed()) return nullptr;
    CreateChunksOfNodes(net);

    // phase 1: each partition contracts random pairs whose result grows by at most one leg
    auto contractPartition = [&net, &randomGenerator](std::shared_ptr<std::vector<std::shared_ptr<Node>>> part) {
        const int threshold = 1;
        int failures = 0;
        while (part->size() != 1 && failures < std::pow(part->size(), 2)) {
            std::uniform_int_distribution<> pick(0, static_cast<int>(part->size()) - 1);
            const int one = pick(randomGenerator), two = pick(randomGenerator);
            if (one == two) continue;
            std::shared_ptr<Node> a = part->at(one), b = part->at(two);
            if (totTimer.getElapsed() > maxTime) return;
            std::shared_ptr<Node> c = net->ContractNodes(a, b, threshold);
            if (c != nullptr) {
                detail::dropPair(*part, one, two);
                failures = 0;
                part->push_back(c);
            } else if (!net->IsDone()) {
                ++failures;
            } else {
                break;
            }
        }
    };
    // The reference runs the partitions on two std::threads sharing one generator (ContractionTools.h:360-368);
    // with the arithmetic asynchronous on the GPU there is nothing to overlap, so they run one after the other --
    // one legal interleaving of the threaded search, and reproducible under SetSeed().
    std::vector<std::shared_ptr<std::vector<std::shared_ptr<Node>>>> parts;
    for (size_t i = 0; i < mPartitionedNodes.size(); ++i) {
        parts.push_back(mPartitionedNodes[i]);
        contractPartition(parts.back());
    }

    // phase 2: everything that is left, starting with rank-reducing contractions only and relaxing the
    // threshold after size^2 consecutive rejections
    std::vector<std::shared_ptr<Node>> left;
    for (const auto &p : parts) left.insert(left.end(), p->begin(), p->end());
    {
        int threshold = -1, fails = 0;
        while (!net->IsDone()) {
            if (fails > std::pow(left.size(), 2)) {
                if (!detail::quietMode()) std::cout << "Nodes left: " << left.size() << std::endl;
                ++threshold;
                fails = 0;
            }
            std::uniform_int_distribution<> pick(0, static_cast<int>(left.size()) - 1);
            const int one = pick(randomGenerator), two = pick(randomGenerator);
            if (one == two || left.at(one)->mContracted || left.at(two)->mContracted) continue;
            std::shared_ptr<Node> a = left.at(one), b = left.at(two);
            if (totTimer.getElapsed() > maxTime) break;
            std::shared_ptr<Node> c = net->ContractNodes(a, b, threshold);
            if (c != nullptr) {
                detail::dropPair(left, one, two);
                left.push_back(c);
                fails = 0;
                threshold = -1;
            } else {
                ++fails;
            }
        }
    }
    if (!net->IsDone()) {
        std::cout << "Error contracting network did not result in a final value..." << std::endl;
        throw ContractionFailure();
    }
    mFinalVal = net->GetFinalValue();
    return net;
}

inline std::shared_ptr<Network> ContractionTools::ContractFromEdges(std::mt19937 &randomGenerator) {
    std::shared_ptr<Network> net = OpenNetwork();
    if (net->HasFailed()) return nullptr;
    net->MoveInitialStatesToBack();
    const std::vector<std::shared_ptr<Node>> &all = net->GetUncontractedNodes();
    const size_t nEdge = 2 * static_cast<size_t>(net->GetNumQubits());
    std::vector<std::shared_ptr<Node>> inner(all.begin(), all.end() - nEdge), working(all.end() - nEdge, all.end());
    int threshold = -1, fails = 0;
    while (!net->IsDone()) {
        if (fails > 100000) { ++threshold; fails = 0; }
        std::uniform_int_distribution<> pickAny(0, static_cast<int>(inner.size() + working.size()) - 1);
        std::uniform_int_distribution<> pickEdge(0, static_cast<int>(working.size()) - 1);
        int one = pickAny(randomGenerator);
        const int two = pickEdge(randomGenerator);
        const bool fromWorking = one >= static_cast<int>(inner.size());
        if (fromWorking) one -= static_cast<int>(inner.size());
        if (fromWorking && one == two) continue;
        std::shared_ptr<Node> a = fromWorking ? working.at(one) : inner.at(one), b = working.at(two);
        std::shared_ptr<Node> c = net->ContractNodes(a, b, threshold);
        if (c == nullptr) { ++fails; continue; }
        if (fromWorking) {
            const int hi = std::max(one, two), lo = std::min(one, two);
            working[hi] = working.back(); working.pop_back();
            working[lo] = working.back(); working.pop_back();
        } else {
            working[two] = working.back(); working.pop_back();
            inner[one] = inner.back(); inner.pop_back();
        }
        working.push_back(c);
        fails = 0;
        threshold = -1;
    }
    if (net->IsDone()) mFinalVal = net->GetFinalValue();
    else std::cout << "Error contracting network did not result in a final value..." << std::endl;
    return net;
}

// Replay a recorded plan: pairs of node ids in mAllNodes numbering (the list of mCreatedFrom pairs).
inline std::shared_ptr<Network> ContractionTools::ContractGivenSequence(const std::vector<std::pair<int, int>> &sequence) {
    if (!mCopyCreated) {
        mNetwork = std::make_shared<Network>(mString, mMeasureFile);
        mNetwork->SetNumThreads(mNumThreadsInNetwork);
    }
    if (mNetwork->HasFailed()) return nullptr;
    for (const auto &step : sequence)
        mNetwork->ContractNodes(mNetwork->GetAllNodes()[step.first], mNetwork->GetAllNodes()[step.second], 100);
    if (!mNetwork->IsDone()) throw ContractionFailure();
    mFinalVal = mNetwork->GetFinalValue();
    return mNetwork;
}

// User file: one "idA idB" pair per line in ORIGINAL node numbering; ids are forwarded to whatever node the
// original one has been merged into (reference ContractionTools.h:188-206).
inline std::shared_ptr<Network> ContractionTools::ContractUserDefinedSequenceOfWires(const std::string &userInputFilePath) {
    std::shared_ptr<Network> net = OpenNetwork();
    if (net->HasFailed()) return nullptr;
    std::ifstream in(userInputFilePath);
    if (!in) throw InvalidFile();
    std::vector<std::pair<int, int>> order;
    const int nNodes = static_cast<int>(net->GetAllNodes().size());
    while (!in.eof()) {
        std::string line;
        std::getline(in, line);
        std::stringstream ss(line);
        int a = 0, b = 0;
        ss >> a >> b;
        if (a < 0 || a > nNodes || b < 0 || b > nNodes) throw InvalidFileFormat();
        order.push_back({a, b});
    }
    std::vector<int> current(nNodes);
    for (int i = 0; i < nNodes; ++i) current[i] = i;
    for (const auto &step : order) {
        const auto &all = net->GetAllNodes();
        if (all[current[step.first]]->mContracted || all[current[step.second]]->mContracted) continue;
        net->ContractNodes(all[current[step.first]], all[current[step.second]], 10000);
        const int newest = static_cast<int>(net->GetAllNodes().size()) - 1;
        current[step.first] = newest;
        current[step.second] = newest;
    }
    if (!net->IsDone()) {
        std::cout << "Error - contraction sequence was incomplete." << std::endl;
        throw InvalidUserContractionSequence();
    }
    mFinalVal = net->GetFinalValue();
    return net;
}

inline std::shared_ptr<Network> ContractionTools::ReduceAndPrintCircuitToTWGraph(const std::string &toPrintTo) const {
    std::shared_ptr<Network> net = OpenNetwork();
    net->ReduceCircuit();
    net->OutputCircuitToTreewidthGraph(toPrintTo);
    return net;
}

inline std::shared_ptr<Network> ContractionTools::ReduceAndPrintCircuitToVisualGraph(const std::string &toPrintTo) const {
    std::shared_ptr<Network> net = OpenNetwork();
    net->ReduceCircuit();
    net->OutputCircuitToVisualGraph(toPrintTo);
    return net;
}

inline const int ContractionTools::CalculateTreewidth(const int qbbseconds, const bool sixtyFourBitOpSystem) const {
    std::shared_ptr<Network> net = OpenNetwork();
    LineGraph lg(net);
    Timer t;
    t.start();
    lg.runQuickBB(qbbseconds, &t, sixtyFourBitOpSystem);
    std::cout << "Please check output/qbb.out for more treewidth and quickbb stats" << std::endl;
    std::ifstream input("output/qbb.out");
    std::string line;
    std::getline(input, line);
    const std::string key = " The treewidth of the graph in the file ";
    const size_t at = line.find(key);
    if (at == std::string::npos) return -1;
    std::stringstream ss(line.substr(at + key.size()));
    std::string skip;
    std::getline(ss, skip, ' ');
    std::getline(ss, skip, ' ');
    int tw = -1;
    ss >> tw;
    return tw;
}

}  // namespace qtorch
