// qtorch_b200/host/preprocess.h -- find a stochastic plan that finishes under a wall-clock cap and return it
// as the list of mCreatedFrom pairs (interface of /root/reference/src/preprocess.h:28-51).  Arms the global
// watchdog (totTimer / maxTime) that the planners and ContractIndices consult between steps.
#pragma once
#include <cstdio>
#include "ContractionTools.h"

namespace qtorch {

inline bool preProcess(const std::string &fileName, std::vector<std::pair<int, int>> &optimalContractionSequence,
                       const double timeThreshold) {
    maxTime = timeThreshold;
    for (int attempt = 0; attempt < 100; ++attempt) {
        totTimer = Timer();
        totTimer.start();
        ContractionTools tools(fileName, "measureTest.txt");
        std::shared_ptr<Network> net = tools.Contract(Stochastic);
        std::remove("measureTest.txt");
        if (totTimer.getElapsed() <= timeThreshold) {
            for (const auto &node : net->GetAllNodes())
                if (!(node->mCreatedFrom.first == 0 && node->mCreatedFrom.second == 0))
                    optimalContractionSequence.push_back(node->mCreatedFrom);
            totTimer.reset();
            return true;
        }
    }
    return false;
}

}  // namespace qtorch
