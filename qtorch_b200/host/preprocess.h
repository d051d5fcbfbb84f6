// qtorch_b200/host/preprocess.h -- search for a stochastic plan that finishes under a wall-clock cap.
// Interface of /root/reference/src/preprocess.h:28-51: preProcess(circuit, sequence out, seconds) -> found?
// Each attempt arms the global watchdog (totTimer / maxTime) that the planners and ContractIndices consult between
// steps, runs ContractionTools::Contract(Stochastic) against "measureTest.txt" (written by the caller, removed here)
// and, if it came in under the cap, hands back the plan as the list of mCreatedFrom pairs in node order -- the form
// ContractGivenSequence replays.
#pragma once
#include <cstdio>
#include <utility>
#include <vector>
#include "ContractionTools.h"

namespace qtorch {
namespace detail {

// the executed plan of a contracted network: (a, b) of every node that was made by a contraction, in id order
inline void appendCreatedFromPairs(const Network &net, std::vector<std::pair<int, int>> &out) {
    for (const std::shared_ptr<Node> &node : net.GetAllNodes()) {
        const std::pair<int, int> &made = node->mCreatedFrom;
        if (made.first != 0 || made.second != 0) out.push_back(made);
    }
}

}  // namespace detail

inline bool preProcess(const std::string &fileName, std::vector<std::pair<int, int>> &optimalContractionSequence,
                       const double timeThreshold) {
    const int kAttempts = 100;
    maxTime = timeThreshold;
    for (int attempt = 0; attempt < kAttempts; ++attempt) {
        totTimer = Timer();
        totTimer.start();
        std::shared_ptr<Network> net;
        {
            ContractionTools search(fileName, "measureTest.txt");
            net = search.Contract(Stochastic);
        }
        std::remove("measureTest.txt");
        if (totTimer.getElapsed() > timeThreshold) continue;          // too slow: draw another plan
        detail::appendCreatedFromPairs(*net, optimalContractionSequence);
        totTimer.reset();
        return true;
    }
    return false;
}

}  // namespace qtorch
