// qtorch_b200/host/Slicing.h -- index-slicing planner and slice dispatcher of the C++14 host.
//
// The reference has no slicing: its only scale knobs are threads and ordering quality, and a
// network whose largest tensor does not fit simply fails.  Fixing the value d in {0,1,2,3} of s wires turns every tensor
// that carries such a wire into its d-slice (rank - 1), leaves the plan's pairwise steps untouched -- the unsliced plan
// stays the reference's plan, bit for bit (Network::ContractNodes, /root/reference/src/Network.h:715-864) -- and the
// network value becomes the sum of the 4^s sliced values.  Steps that no cut wire reaches are the same in every slice:
// HoistInvariant moves them to the front (a stable re-ordering of independent steps) and the device runs them once per
// amplitude.  Slices are dealt round-robin over the ranks of a job (device::Job) and the partial sums meet in ONE
// in-stream ncclAllReduce per amplitude -- the reduction that replaces `f_pVal +=` (/root/reference/src/maxcut.cpp:196).
//
// A plan here is the reference's own record of a contraction -- mCreatedFrom pairs plus leg maps (Network::GetPlan()) --
// with tensor ids 0..n-1 for the original nodes and n+i for the result of step i.
// qtorch_b200/slicing.py is the same planner in Python (used by the CPU tests with the oracle); tests/test_slicing_host.py
// checks that both choose the same wires and emit the same sliced plans.
#pragma once

#include <algorithm>
#include <complex>
#include <cstring>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

#include "DeviceEngine.h"
#include "Network.h"

namespace qtorch {
namespace slicing {

struct Step {
    int a, b;
    std::vector<int> posA, posB;
};
struct Plan {
    std::vector<int> inputRanks;
    std::vector<Step> steps;
};
typedef unsigned __int128 Units;          // sum of 4^(rC+k): exact up to rank-16 steps with 16 shared legs

// every step a Network executed, as a plan over its original nodes
inline Plan PlanOfNetwork(const Network &net) {
    Plan p;
    for (int i = 0; i < net.GetNumOriginalNodes(); ++i) p.inputRanks.push_back(net.GetAllNodes()[i]->mRank);
    for (const PlanRecord &r : net.GetPlan()) p.steps.push_back({r.a, r.b, r.posA, r.posB});
    return p;
}

// Wire label of every leg of every tensor (legs contracted together are the same wire).  A label is the union-find root
// of its legs, encoded as tensor * 32 + leg of an ORIGINAL node (ranks <= 16 < 32), so labels order like (tensor, leg).
struct Labels {
    std::vector<std::vector<int>> legs;                                  // per tensor id (inputs, then step results)
    std::map<int, std::vector<std::pair<int, int>>> ends;                // label -> (input tensor, leg) of its ends
};

inline Labels LabelWires(const Plan &plan) {
    const int n = static_cast<int>(plan.inputRanks.size());
    std::unordered_map<int, int> parent;
    auto find = [&parent](int x) {
        for (;;) {
            auto it = parent.find(x);
            if (it == parent.end() || it->second == x) return x;
            x = it->second;
        }
    };
    std::vector<std::vector<int>> raw;                                   // un-resolved leg names
    for (int t = 0; t < n; ++t) {
        std::vector<int> l;
        for (int j = 0; j < plan.inputRanks[t]; ++j) l.push_back(t * 32 + j);
        raw.push_back(l);
    }
    for (const Step &s : plan.steps) {
        const std::vector<int> la = raw[s.a], lb = raw[s.b];
        for (size_t j = 0; j < s.posA.size(); ++j) {
            const int rb = find(lb[s.posB[j]]), ra = find(la[s.posA[j]]);
            parent[rb] = ra;                                             // B's end joins A's (same direction as slicing.py)
        }
        std::vector<int> freeLegs;
        for (size_t i = 0; i < la.size(); ++i) if (std::find(s.posA.begin(), s.posA.end(), static_cast<int>(i)) == s.posA.end()) freeLegs.push_back(la[i]);
        for (size_t i = 0; i < lb.size(); ++i) if (std::find(s.posB.begin(), s.posB.end(), static_cast<int>(i)) == s.posB.end()) freeLegs.push_back(lb[i]);
        raw.push_back(freeLegs);
    }
    Labels out;
    for (const auto &t : raw) {
        std::vector<int> l;
        for (int x : t) l.push_back(find(x));
        out.legs.push_back(l);
    }
    for (int t = 0; t < n; ++t)
        for (int j = 0; j < plan.inputRanks[t]; ++j) out.ends[find(t * 32 + j)].push_back({t, j});
    return out;
}

// (units, peak rank) of ONE slice of the plan with the wires in `removed` fixed
inline std::pair<Units, int> PlanCost(const Plan &plan, const Labels &lab, const std::set<int> &removed) {
    const int n = static_cast<int>(plan.inputRanks.size());
    Units units = 0;
    int peak = 0;
    auto live = [&removed](const std::vector<int> &l) {
        int c = 0;
        for (int x : l) if (!removed.count(x)) ++c;
        return c;
    };
    for (size_t i = 0; i < plan.steps.size(); ++i) {
        const Step &s = plan.steps[i];
        const std::vector<int> &la = lab.legs[s.a];
        int k = 0;
        for (int x : s.posA) if (!removed.count(la[x])) ++k;
        const int rc = live(lab.legs[n + i]);
        units += static_cast<Units>(1) << (2 * (rc + k));
        peak = std::max(peak, std::max(rc, std::max(live(la), live(lab.legs[s.b]))));
    }
    return {units, peak};
}

// greedy: repeatedly cut the wire that minimises (peak rank, units) of the remaining plan; candidates are the internal
// wires (two input ends) of the currently largest tensors; ties go to the smallest label
inline std::vector<int> ChooseWires(const Plan &plan, int nSliceWires) {
    const Labels lab = LabelWires(plan);
    std::vector<int> removed;
    for (int round = 0; round < nSliceWires; ++round) {
        const std::set<int> gone(removed.begin(), removed.end());
        const int peak = PlanCost(plan, lab, gone).second;
        std::set<int> cands;
        for (const auto &t : lab.legs) {
            std::vector<int> liveLegs;
            for (int x : t) if (!gone.count(x)) liveLegs.push_back(x);
            if (static_cast<int>(liveLegs.size()) != peak) continue;
            for (int x : liveLegs) {
                auto it = lab.ends.find(x);
                if (it != lab.ends.end() && it->second.size() == 2) cands.insert(x);
            }
        }
        if (cands.empty()) break;
        bool have = false;
        int best = 0;
        std::pair<int, Units> bestKey{0, 0};
        for (int w : cands) {                                            // ascending label order
            std::set<int> trial = gone;
            trial.insert(w);
            const auto c = PlanCost(plan, lab, trial);
            const std::pair<int, Units> key{c.second, c.first};
            if (!have || key < bestKey) { have = true; best = w; bestKey = key; }
        }
        removed.push_back(best);
    }
    return removed;
}

struct SlicedPlan {
    Plan plan;                                                   // ONE slice: same steps, cut legs dropped, invariant steps first
    int nInvariant = 0;                                          // leading steps that no cut wire reaches
    std::vector<int> wires;                                      // the cut wires (labels)
    std::map<int, std::vector<std::pair<int, int>>> cuts;        // input tensor -> sorted (leg, wire) pairs that are fixed
    Units unitsPerSlice = 0, unitsInvariant = 0;
    int peakRank = 0;
    size_t NumSlices() const { return static_cast<size_t>(1) << (2 * wires.size()); }
    // digit of wire i in slice u: u = sum_i digit_i * 4^(s-1-i)  (wire 0 most significant, like itertools.product)
    int Digit(size_t slice, size_t wire) const { return static_cast<int>((slice >> (2 * (wires.size() - 1 - wire))) & 3u); }
};

// stable re-ordering: steps that do not depend (transitively) on a tensor in `variant` first
inline int HoistInvariant(Plan &plan, const std::set<int> &variant) {
    const int n = static_cast<int>(plan.inputRanks.size());
    std::set<int> dep(variant.begin(), variant.end());
    std::vector<int> inv, var;
    for (size_t i = 0; i < plan.steps.size(); ++i) {
        if (dep.count(plan.steps[i].a) || dep.count(plan.steps[i].b)) { dep.insert(n + static_cast<int>(i)); var.push_back(static_cast<int>(i)); }
        else inv.push_back(static_cast<int>(i));
    }
    std::vector<int> order = inv;
    order.insert(order.end(), var.begin(), var.end());
    std::vector<int> newId(n + plan.steps.size());
    for (int t = 0; t < n; ++t) newId[t] = t;
    for (size_t pos = 0; pos < order.size(); ++pos) newId[n + order[pos]] = n + static_cast<int>(pos);
    std::vector<Step> out;
    for (int i : order) out.push_back({newId[plan.steps[i].a], newId[plan.steps[i].b], plan.steps[i].posA, plan.steps[i].posB});
    plan.steps = out;
    return static_cast<int>(inv.size());
}

inline SlicedPlan SlicePlan(const Plan &plan, const std::vector<int> &wires) {
    const Labels lab = LabelWires(plan);
    const std::set<int> gone(wires.begin(), wires.end());
    SlicedPlan sp;
    sp.wires = wires;
    for (int w : wires) {
        auto it = lab.ends.find(w);
        if (it == lab.ends.end() || it->second.size() != 2) throw InvalidFunctionInput();      // not an internal wire of the plan
        for (const auto &e : it->second) sp.cuts[e.first].push_back({e.second, w});
    }
    for (auto &c : sp.cuts) std::sort(c.second.begin(), c.second.end());
    sp.plan.inputRanks = plan.inputRanks;
    for (const auto &c : sp.cuts) sp.plan.inputRanks[c.first] -= static_cast<int>(c.second.size());
    for (const Step &s : plan.steps) {
        const std::vector<int> &la = lab.legs[s.a], &lb = lab.legs[s.b];
        auto newPos = [&gone](const std::vector<int> &l, int i) {
            int c = 0;
            for (int j = 0; j < i; ++j) if (!gone.count(l[j])) ++c;
            return c;
        };
        Step t{s.a, s.b, {}, {}};
        for (size_t j = 0; j < s.posA.size(); ++j) {
            if (gone.count(la[s.posA[j]])) continue;
            t.posA.push_back(newPos(la, s.posA[j]));
            t.posB.push_back(newPos(lb, s.posB[j]));
        }
        sp.plan.steps.push_back(t);
    }
    std::set<int> variant;
    for (const auto &c : sp.cuts) variant.insert(c.first);
    sp.nInvariant = HoistInvariant(sp.plan, variant);
    // cost of one slice and of the shared prefix
    std::vector<int> rk = sp.plan.inputRanks;
    for (size_t i = 0; i < sp.plan.steps.size(); ++i) {
        const Step &s = sp.plan.steps[i];
        const int k = static_cast<int>(s.posA.size()), rc = rk[s.a] + rk[s.b] - 2 * k;
        rk.push_back(rc);
        const Units u = static_cast<Units>(1) << (2 * (rc + k));
        sp.unitsPerSlice += u;
        if (static_cast<int>(i) < sp.nInvariant) sp.unitsInvariant += u;
        sp.peakRank = std::max(sp.peakRank, std::max(rc, std::max(rk[s.a], rk[s.b])));
    }
    return sp;
}

// the d-slice of one input tensor: legs listed in `cut` (leg, digit) are fixed, the others keep their order
inline std::vector<std::complex<double>> SliceTensor(const std::vector<std::complex<double>> &full, int rank, const std::vector<std::pair<int, int>> &cut) {
    const int newRank = rank - static_cast<int>(cut.size());
    std::vector<int> keep;
    size_t base = 0;
    for (int l = 0; l < rank; ++l) {
        bool fixed = false;
        for (const auto &c : cut) if (c.first == l) { base += static_cast<size_t>(c.second) << (2 * l); fixed = true; }
        if (!fixed) keep.push_back(l);
    }
    std::vector<std::complex<double>> out(static_cast<size_t>(1) << (2 * newRank));
    for (size_t e = 0; e < out.size(); ++e) {
        size_t off = base;
        for (int j = 0; j < newRank; ++j) off += ((e >> (2 * j)) & 3u) << (2 * keep[j]);
        out[e] = full[off];
    }
    return out;
}

// input tensors of slice `slice` (digits of the cut wires as in SlicedPlan::Digit)
inline std::vector<std::vector<std::complex<double>>> SliceInputs(const SlicedPlan &sp, const std::vector<int> &fullRanks,
                                                                  const std::vector<std::vector<std::complex<double>>> &inputs, size_t slice) {
    std::vector<std::vector<std::complex<double>>> out;
    for (size_t t = 0; t < inputs.size(); ++t) {
        auto it = sp.cuts.find(static_cast<int>(t));
        if (it == sp.cuts.end()) { out.push_back(inputs[t]); continue; }
        std::vector<std::pair<int, int>> cut;
        for (const auto &lw : it->second) {
            const size_t wi = std::find(sp.wires.begin(), sp.wires.end(), lw.second) - sp.wires.begin();
            cut.push_back({lw.first, sp.Digit(slice, wi)});
        }
        out.push_back(SliceTensor(inputs[t], fullRanks[t], cut));
    }
    return out;
}

inline std::vector<qtb_plan_step> ToAbiSteps(const Plan &plan) {
    std::vector<qtb_plan_step> out;
    for (const Step &s : plan.steps) {
        qtb_plan_step st;
        std::memset(&st, 0, sizeof(st));
        st.a = s.a; st.b = s.b; st.k = static_cast<int>(s.posA.size());
        for (int j = 0; j < st.k; ++j) { st.pos_a[j] = static_cast<int8_t>(s.posA[j]); st.pos_b[j] = static_cast<int8_t>(s.posB[j]); }
        out.push_back(st);
    }
    return out;
}

}  // namespace slicing

// One network, index-sliced over the ranks of a job.  Construct once per topology; every amplitude (a new set of input
// tensors: other measurement caps, other angles) is staged, begun and read back without a host synchronisation in between,
// and up to two amplitudes may be in flight (Begin(i+1) before End(i)), which hides the run-once prefix of one behind the
// slices of the other.
class SlicedContraction {
public:
    typedef std::vector<std::vector<std::complex<double>>> Tensors;

    // nSliceWires < 0: as few wires as give every rank a slice.  lanes: plan replicas per rank (see qtb_sliced_create).
    SlicedContraction(const slicing::Plan &plan, int nSliceWires, const device::Job &job, int lanes = 2, bool reduceOverRanks = true)
        : mFullRanks(plan.inputRanks), mRank(job.rank), mWorld(job.world), mReduce(reduceOverRanks && job.world > 1) {
        if (nSliceWires < 0) { nSliceWires = 0; while ((1 << (2 * nSliceWires)) < job.world) ++nSliceWires; }
        mSliced = slicing::SlicePlan(plan, slicing::ChooseWires(plan, nSliceWires));
        for (size_t u = static_cast<size_t>(mRank); u < mSliced.NumSlices(); u += static_cast<size_t>(mWorld)) mOwned.push_back(u);
        const std::vector<qtb_plan_step> steps = slicing::ToAbiSteps(mSliced.plan);
        device::check(qtb_sliced_create(device::Engine::Get().ctx(), static_cast<int>(mSliced.plan.inputRanks.size()), mSliced.plan.inputRanks.data(),
                                        static_cast<int>(steps.size()), steps.data(), mSliced.nInvariant, std::max(1, std::min<int>(lanes, std::max<size_t>(mOwned.size(), 1))), &mHandle));
    }
    ~SlicedContraction() {
        if (mHandle && device::Engine::Get().alive()) {
            for (auto &p : mPending) if (p) qtb_read_scalar_end(device::Engine::Get().ctx(), p, nullptr);
            qtb_sliced_destroy(device::Engine::Get().ctx(), mHandle);
        }
    }
    SlicedContraction(const SlicedContraction &) = delete;
    SlicedContraction &operator=(const SlicedContraction &) = delete;

    const slicing::SlicedPlan &Sliced() const { return mSliced; }
    const std::vector<size_t> &OwnedSlices() const { return mOwned; }
    int LaunchesPerSlice(int *prefixLaunches = nullptr) const { return qtb_sliced_launches(mHandle, prefixLaunches); }

    // host -> device: the owned slices of this amplitude's input tensors into slot bank `bank` (0 or 1).  Alternate the
    // banks between consecutive amplitudes when two are kept in flight.
    void Stage(const Tensors &inputs, int bank = 0) {
        qtb_ctx *ctx = device::Engine::Get().ctx();
        for (size_t j = 0; j < mOwned.size(); ++j) {
            const Tensors sl = slicing::SliceInputs(mSliced, mFullRanks, inputs, mOwned[j]);
            std::vector<const double *> ptrs;
            for (const auto &t : sl) ptrs.push_back(reinterpret_cast<const double *>(t.data()));
            device::check(qtb_sliced_stage(ctx, mHandle, static_cast<int>(bank * mOwned.size() + j), ptrs.data()));
        }
    }
    // enqueue one amplitude on the slices staged in `bank`; returns a ticket for End()
    int Begin(int bank = 0) {
        std::vector<int> slots;
        for (size_t j = 0; j < mOwned.size(); ++j) slots.push_back(static_cast<int>(bank * mOwned.size() + j));
        qtb_scalar_read *rd = nullptr;
        device::check(qtb_sliced_begin(device::Engine::Get().ctx(), mHandle, slots.data(), static_cast<int>(slots.size()), mReduce ? 1 : 0, &rd));
        for (size_t i = 0; i < mPending.size(); ++i) if (!mPending[i]) { mPending[i] = rd; return static_cast<int>(i); }
        mPending.push_back(rd);
        return static_cast<int>(mPending.size()) - 1;
    }
    // the network value: sum over all slices and ranks
    std::complex<double> End(int ticket) {
        double v[2] = {0.0, 0.0};
        qtb_scalar_read *rd = mPending.at(static_cast<size_t>(ticket));
        mPending[static_cast<size_t>(ticket)] = nullptr;
        device::check(qtb_read_scalar_end(device::Engine::Get().ctx(), rd, v));
        return {v[0], v[1]};
    }
    std::complex<double> Contract(const Tensors &inputs) {
        Stage(inputs, 0);
        return End(Begin(0));
    }

private:
    std::vector<int> mFullRanks;
    int mRank, mWorld;
    bool mReduce;
    slicing::SlicedPlan mSliced;
    std::vector<size_t> mOwned;
    qtb_sliced *mHandle{nullptr};
    std::vector<qtb_scalar_read *> mPending;
};

}  // namespace qtorch
