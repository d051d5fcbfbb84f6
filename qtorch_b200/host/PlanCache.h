// qtorch_b200/host/PlanCache.h -- compiled circuits and the plan cache.
//
// The contraction plan of a network depends on the circuit's topology and on the ordering, not on gate angles and not on
// the measurement string: measurement / trace nodes are rank-1 caps on every qubit line whatever they measure
// (/root/reference/src/Network.h:199-237), and ReduceCircuit / LGContract decide by ranks and wires alone
// (src/Network.h:1015-1118, src/LineGraph.h:306-399).  The reference nevertheless re-parses, re-reduces and re-walks the
// ordering for every evaluation (src/main.cpp:74-198; src/maxcut.cpp:57-115 does it once per qubit).  A CompiledCircuit
// does that host work ONCE -- in plan-only mode, no arithmetic -- compiles the recorded steps into a device plan (CUDA
// graph), and then evaluates any number of measurement strings by swapping the n cap tensors: one small H2D, one graph
// launch, one 16-byte D2H.  PlanCache keys compiled circuits on (qasm file, ordering file, reduce) plus the files' size and
// modification time, so an unchanged pair is compiled once per process.
#pragma once

#include <sys/stat.h>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "LineGraph.h"
#include "Network.h"

namespace qtorch {

class CompiledCircuit {
public:
    // qbbOut empty: in-process min-fill ordering (LineGraph::runMinFill), no file involved
    CompiledCircuit(const std::string &qasm, const std::string &qbbOut, bool reduce) {
        const bool before = device::Engine::PlanOnly();
        device::Engine::SetPlanOnly(true);
        std::vector<qtb_plan_step> steps;
        try {
            std::shared_ptr<Network> net = std::make_shared<Network>(qasm, std::string("/dev/null"));       // every qubit traced: caps are placeholders
            mNumQubits = net->GetNumQubits();
            for (int i = 0; i < net->GetNumOriginalNodes(); ++i) mInputs.push_back(net->GetAllNodes()[i]->GetTensorVals());
            if (reduce) net->ReduceCircuit();
            LineGraph lg(net);
            lg.SetQBBOutFiles("/dev/null", qbbOut, "/dev/null");
            if (qbbOut.empty()) lg.runMinFill();
            mOk = lg.LGContract();
            mUnits = net->getNumFloatOps();
            mNodes = static_cast<int>(net->GetAllNodes().size());
            for (const PlanRecord &r : net->GetPlan()) {
                qtb_plan_step s;
                std::memset(&s, 0, sizeof(s));
                s.a = r.a; s.b = r.b; s.k = static_cast<int>(r.posA.size());
                for (int j = 0; j < s.k; ++j) { s.pos_a[j] = static_cast<int8_t>(r.posA[j]); s.pos_b[j] = static_cast<int8_t>(r.posB[j]); }
                steps.push_back(s);
            }
        } catch (...) {
            device::Engine::SetPlanOnly(before);
            throw;
        }
        device::Engine::SetPlanOnly(before);
        std::vector<int> ranks;
        for (const auto &in : mInputs) { int r = 0; while ((static_cast<size_t>(1) << (2 * r)) < in.size()) ++r; ranks.push_back(r); }
        device::check(qtb_plan_create(device::Engine::Get().ctx(), static_cast<int>(ranks.size()), ranks.data(), static_cast<int>(steps.size()), steps.data(), &mPlan));
        mFirstCap = mInputs.size() - static_cast<size_t>(mNumQubits);          // the n caps are the last original nodes (Network.h:236)
    }
    ~CompiledCircuit() {
        if (mPlan && device::Engine::Get().alive()) qtb_plan_destroy(device::Engine::Get().ctx(), mPlan);
    }
    CompiledCircuit(const CompiledCircuit &) = delete;
    CompiledCircuit &operator=(const CompiledCircuit &) = delete;

    // the network value for a measurement string ("0 1 T X ...": one character per qubit, whitespace ignored, missing
    // entries trace the qubit out -- the reference's measurement-file grammar)
    std::complex<double> Evaluate(const std::string &measurementText) {
        std::istringstream in(measurementText);
        for (int q = 0; q < mNumQubits; ++q) {
            char c;
            const char m = (in >> c) ? c : 'T';
            mInputs[mFirstCap + static_cast<size_t>(q)] = Network::MakeMeasurementCap(m)->GetTensorVals();
        }
        std::vector<const double *> ptrs;
        for (const auto &t : mInputs) ptrs.push_back(reinterpret_cast<const double *>(t.data()));
        double out[2] = {0.0, 0.0};
        device::check(qtb_plan_run_host(device::Engine::Get().ctx(), mPlan, ptrs.data(), out));
        ++mEvaluations;
        return {out[0], out[1]};
    }
    std::complex<double> EvaluateFile(const std::string &measurementFile) {
        std::ifstream f(measurementFile);
        std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        if (!f.is_open()) std::cout << "Measurement file failed to open - all qubits will be traced out" << std::endl;
        return Evaluate(text);
    }

    bool Ok() const { return mOk; }
    long long Units() const { return mUnits; }                    // the reference's getNumFloatOps() for this plan
    int NumNodes() const { return mNodes; }
    int NumQubits() const { return mNumQubits; }
    long long Evaluations() const { return mEvaluations; }
    int Launches() const { return qtb_plan_launches(mPlan); }

private:
    qtb_plan *mPlan{nullptr};
    std::vector<std::vector<std::complex<double>>> mInputs;
    size_t mFirstCap{0};
    int mNumQubits{0}, mNodes{0};
    long long mUnits{0}, mEvaluations{0};
    bool mOk{false};
};

class PlanCache {
public:
    static PlanCache &Get() {
        static PlanCache c;
        return c;
    }
    // the compiled circuit of (qasm, ordering, reduce); compiled on first use, re-compiled when a file changed on disk
    std::shared_ptr<CompiledCircuit> Lookup(const std::string &qasm, const std::string &qbbOut, bool reduce) {
        const std::string key = Stamp(qasm) + "|" + (qbbOut.empty() ? std::string("<minfill>") : Stamp(qbbOut)) + "|" + (reduce ? "1" : "0");
        auto it = mEntries.find(key);
        if (it != mEntries.end()) { ++mHits; return it->second; }
        ++mMisses;
        std::shared_ptr<CompiledCircuit> c = std::make_shared<CompiledCircuit>(qasm, qbbOut, reduce);
        mEntries[key] = c;
        return c;
    }
    void Clear() { mEntries.clear(); }
    long long Hits() const { return mHits; }
    long long Misses() const { return mMisses; }

private:
    static std::string Stamp(const std::string &path) {
        struct stat st;
        std::ostringstream os;
        os << path;
        if (stat(path.c_str(), &st) == 0) os << ":" << static_cast<long long>(st.st_size) << ":" << static_cast<long long>(st.st_mtim.tv_sec) << "." << st.st_mtim.tv_nsec;
        return os.str();
    }
    std::map<std::string, std::shared_ptr<CompiledCircuit>> mEntries;
    long long mHits{0}, mMisses{0};
};

}  // namespace qtorch
