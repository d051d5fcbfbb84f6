// qtorch_b200/host/Wire.h -- one leg of the tensor network (dimension 4, density-matrix picture).
// Interface of /root/reference/src/Wire.h:26-60: two weak endpoints (NodeA = input side, NodeB = output
// side), the qubit line it sits on, a line-graph id and the "already summed" flag LGContract reads.
#pragma once
#include <memory>
#include "Exceptions.h"

namespace qtorch {

class Node;

class Wire {
public:
    explicit Wire(std::shared_ptr<Node> nodeA, std::shared_ptr<Node> nodeB, int qubitNum)
        : mEndA(nodeA), mEndB(nodeB), mQubit(qubitNum) {}

    void SetNodeA(std::shared_ptr<Node> n) { mEndA = n; }
    void SetNodeB(std::shared_ptr<Node> n) { mEndB = n; }
    std::weak_ptr<Node> GetNodeA() { return mEndA; }
    std::weak_ptr<Node> GetNodeB() { return mEndB; }

    int GetQubitNumber() { return mQubit; }
    void SetQubitNumber(int q) { mQubit = q; }

    void SetWireID(int id) { mLineGraphId = id; }
    int GetWireID() { return mLineGraphId; }

    void SetIsContracted(bool v) { mSummed = v; }
    bool IsContracted() { return mSummed; }

private:
    std::weak_ptr<Node> mEndA, mEndB;
    int mQubit;
    int mLineGraphId{0};
    bool mSummed{false};
};

}  // namespace qtorch
