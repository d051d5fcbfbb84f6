// qtorch_b200/host/LineGraph.h -- line-graph + QuickBB tree-decomposition ordering, and the contraction
// loop that follows it.  Interface and file formats of /root/reference/src/LineGraph.h:44-410.
//
// The ordering is host work and stays as it was: vertices of L(G) are the wires in order of first appearance
// while walking the uncontracted nodes, edges join every pair of wires meeting at a node (emitted
// (earlier, later) per node), "lg.cnf" is handed to the external quickbb_64 binary, and the wires are then
// contracted in the returned order through Network::ContractNodes -- which is where the B200 engine takes over.
// Freezing qbb.out (qbbonly / readqbbresonly, main.cpp) freezes the plan.
#pragma once

#include <sys/stat.h>
#include <array>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "Exceptions.h"
#include "Network.h"
#include "Timer.h"

namespace qtorch {

class LineGraph {
public:
    LineGraph(std::shared_ptr<Network> origGraph) { Build(origGraph); }

    bool runQuickBB(int MaxTimeInSec, Timer *tim = NULL, bool sixtyFourBit = true);
    void Reset(std::shared_ptr<Network> inpNetwork = nullptr) {
        if (inpNetwork == nullptr) { origNetwork->Reset(); return; }
        Build(inpNetwork);
    }
    bool LGContract();

    void SetQBBOutDirectory(std::string &pathToDirectory) {
        cnfName = pathToDirectory + "lg.cnf";
        qbbOutName = pathToDirectory + "qbb.out";
        qbbStatsName = pathToDirectory + "qbb-stats.out";
    }
    void SetQBBOutFiles(const std::string &cnfNew, const std::string &qbbOutNew, const std::string &qbbStatsNew) {
        cnfName = cnfNew; qbbOutName = qbbOutNew; qbbStatsName = qbbStatsNew;
    }
    // additions: sizes of L(G) and the cnf text, for tests
    size_t NumLineGraphVertices() const { return GraphWires.size(); }
    size_t NumLineGraphEdges() const { return LGEdges.size(); }
    void WriteCnf(std::ostream &os) const {
        os << "p cnf " << GraphWires.size() << " " << LGEdges.size() << std::endl;
        for (const auto &e : LGEdges) os << e[0]->GetWireID() + 1 << " " << e[1]->GetWireID() + 1 << " " << 0 << std::endl;
    }

private:
    void Build(std::shared_ptr<Network> net);

    std::shared_ptr<Network> origNetwork;
    std::vector<std::shared_ptr<Wire>> GraphWires;                    // vertex i of L(G)
    std::vector<std::array<std::shared_ptr<Wire>, 2>> LGEdges;
    std::string cnfName = "output/lg.cnf";
    std::string qbbOutName = "output/qbb.out";
    std::string qbbStatsName = "output/qbb-stats.out";
};

inline void LineGraph::Build(std::shared_ptr<Network> net) {
    origNetwork = net;
    GraphWires.clear();
    LGEdges.clear();
    std::unordered_map<Wire *, int> seen;
    const std::vector<std::shared_ptr<Node>> nodes = net->GetUncontractedNodes();
    for (const auto &node : nodes) {
        const std::vector<std::shared_ptr<Wire>> wires = node->GetWires();
        for (size_t i = 0; i < wires.size(); ++i) {
            if (seen.find(wires[i].get()) == seen.end()) {
                const int id = static_cast<int>(GraphWires.size());
                seen.emplace(wires[i].get(), id);
                wires[i]->SetWireID(id);
                GraphWires.push_back(wires[i]);
            }
            for (size_t j = 0; j < i; ++j) LGEdges.push_back({wires[j], wires[i]});
        }
    }
    if (!detail::quietMode()) std::cout << "GraphWires.size(): " << GraphWires.size() << std::endl;
}

inline bool LineGraph::runQuickBB(int MaxTimeInSec, Timer *tim, bool sixtyFourBit) {
    mkdir("output", 0755);
    std::remove(qbbOutName.c_str());
    {
        std::ofstream cnf(cnfName);
        WriteCnf(cnf);
    }
    std::cout << "===== Output from QuickBB =====" << std::endl;
    std::ostringstream cmd;
    cmd << "quickbb_" << (sixtyFourBit ? "64" : "32") << " --min-fill-ordering --lb --time " << MaxTimeInSec << " --outfile "
        << qbbOutName << " --statfile " << qbbStatsName << " --cnffile " << cnfName;
    std::cout << "Executing:   " << cmd.str() << std::endl;
    const int rc = system(cmd.str().c_str());
    (void)rc;
    std::cout << "===== End of QuickBB output =====" << std::endl << std::endl;
    if (tim) std::cout << "Time elapsed after outputting line graph and running QuickBB: { " << tim->getElapsed() << " }\n";
    return true;
}

inline bool LineGraph::LGContract() {
    std::ifstream fQbb(qbbOutName);
    if (!fQbb) {
        std::cout << "Unable to open qbb file: " << qbbOutName << std::endl;
        throw QbbFailure();
    }
    std::vector<int> order;
    std::string line;
    bool found = false;
    while (std::getline(fQbb, line)) {
        if (line != " The optimal ordering is ") continue;
        std::getline(fQbb, line);
        std::stringstream ss(line);
        for (size_t i = 0; i < GraphWires.size(); ++i) {
            int w = 0;
            ss >> w;
            order.push_back(w);
        }
        found = true;
        break;
    }
    if (!found) {
        std::cout << "ERROR reading quickbb contr ordering.\n";
        return false;
    }
    if (!detail::quietMode()) {
        std::cout << "The contraction ordering read from qbb (should match above output): \n";
        for (int w : order) std::cout << w << " ";
        std::cout << "\n\n";
    }
    // contract wire by wire; a wire already summed by an earlier step (parallel wires) is skipped
    for (int w1 : order) {
        // QuickBB leaves isolated line-graph vertices out of its ordering, so the line can hold fewer numbers than
        // there are wires; the reference then indexes GraphWires[-1] (LineGraph.h:328-360, undefined behaviour that
        // happens to be harmless).  Skip such entries: the untouched components surface below as ContractionFailure,
        // exactly the reference's outcome for disconnected networks.
        if (w1 < 1 || w1 > static_cast<int>(GraphWires.size())) continue;
        std::shared_ptr<Wire> w = GraphWires[w1 - 1];
        if (!w->IsContracted()) origNetwork->ContractNodes(w->GetNodeA().lock(), w->GetNodeB().lock(), 100);
    }
    const std::vector<std::shared_ptr<Node>> &left = origNetwork->GetUncontractedNodes();
    if (left.size() != 1) {
        std::cout << "ERROR. After contraction, there is more than one remaining node.\n";
        throw ContractionFailure();
    }
    if (left[0]->mRank != 0) {
        std::cout << "ERROR. Final node has more than one value.\n";
        throw ContractionFailure();
    }
    if (!detail::quietMode()) std::cout << "Result of contraction:\n" << origNetwork->GetFinalValue() << "\n";
    return true;
}

}  // namespace qtorch
