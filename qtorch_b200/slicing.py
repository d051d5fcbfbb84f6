"""Index-slicing planner (host side): cut a contraction plan into 4^s independent sub-plans with identical step shapes.

The reference has no slicing (its only scale knobs are threads and ordering quality); this is the
B200-side answer to "a network whose largest tensor does not fit, or that should spread over several GPUs": fixing the
value d in {0,1,2,3} of s wires turns every tensor that carries such a wire into its d-slice (rank - 1), leaves the plan's
pairwise steps untouched (so the unsliced plan stays the reference's plan, bit for bit) and the network value becomes the
sum of the 4^s sliced values.  Steps that no cut wire reaches are the same in every slice: hoist_invariant moves them to
the front of the plan and the device runs them once per amplitude instead of once per slice.  All slices share one compiled device plan; only the (small) input tensors differ, so they are
staged as input slots (qtb_plan_stage_inputs) and dealt round-robin to ranks by qtorch_b200.dispatch.

A plan is (input_ranks, steps) with steps = [(a, b, posA, posB)] in the reference's mCreatedFrom numbering.
"""
import itertools

import numpy as np


def label_wires(input_ranks, steps):
    """Give every leg of every tensor a wire label.  Legs contracted together are the same wire.
    Returns (labels per tensor id, {label: [(input tensor, leg), ...]})."""
    n = len(input_ranks)
    parent = {}

    def find(x):
        while parent.get(x, x) != x:
            parent[x] = parent.get(parent[x], parent[x])
            x = parent[x]
        return x

    legs = [[(t, l) for l in range(r)] for t, r in enumerate(input_ranks)]
    tensors = list(legs)
    for (a, b, pa, pb) in steps:
        la, lb = tensors[a], tensors[b]
        for x, y in zip(pa, pb):
            parent[find(lb[y])] = find(la[x])
        free = [l for i, l in enumerate(la) if i not in pa] + [l for i, l in enumerate(lb) if i not in pb]
        tensors.append(free)
    labelled = [[find(l) for l in t] for t in tensors]
    ends = {}
    for t in range(n):
        for l in range(input_ranks[t]):
            ends.setdefault(find((t, l)), []).append((t, l))
    return labelled, ends


def plan_cost(input_ranks, steps, removed=frozenset()):
    """(units, peak rank) of the plan with the wires in `removed` sliced away; units = sum 4^(rC+k) per slice"""
    labelled, _ = label_wires(input_ranks, steps)
    n = len(input_ranks)
    units, peak = 0, 0
    for i, (a, b, pa, pb) in enumerate(steps):
        la, lb = labelled[a], labelled[b]
        k = sum(1 for x in pa if la[x] not in removed)
        rc = sum(1 for l in labelled[n + i] if l not in removed)
        units += 4 ** (rc + k)
        peak = max(peak, rc, sum(1 for l in la if l not in removed), sum(1 for l in lb if l not in removed))
    return units, peak


def choose_wires(input_ranks, steps, n_slice_wires):
    """greedy: repeatedly slice the wire that minimises (peak rank, units) of the remaining plan"""
    labelled, ends = label_wires(input_ranks, steps)
    n = len(input_ranks)
    removed = []
    for _ in range(n_slice_wires):
        _, peak = plan_cost(input_ranks, steps, frozenset(removed))
        # candidates: wires of the largest tensors that are eventually contracted (two input ends)
        cands = set()
        for t in labelled:
            live = [l for l in t if l not in removed]
            if len(live) == peak:
                cands.update(l for l in live if len(ends.get(l, ())) == 2)
        if not cands:
            break
        best = min(sorted(cands), key=lambda w: plan_cost(input_ranks, steps, frozenset(removed + [w]))[::-1])
        removed.append(best)
    return removed


def slice_plan(input_ranks, steps, wires):
    """Plan of ONE slice: same steps, sliced legs dropped.  Returns (ranks', steps', cuts) with
    cuts[t] = sorted legs of input tensor t that are fixed (descending removal order is the caller's business)."""
    labelled, ends = label_wires(input_ranks, steps)
    n = len(input_ranks)
    wires = list(wires)
    gone = set(wires)
    for w in wires:
        if len(ends.get(w, ())) != 2:
            raise ValueError("wire %r is not an internal wire of the plan" % (w,))
    cuts = {}
    for w in wires:
        for (t, l) in ends[w]:
            cuts.setdefault(t, []).append((l, w))
    new_ranks = [r - len(cuts.get(t, ())) for t, r in enumerate(input_ranks)]
    new_steps = []
    for (a, b, pa, pb) in steps:
        la, lb = labelled[a], labelled[b]

        def newpos(lab, i):
            return sum(1 for j in range(i) if lab[j] not in gone)

        npa, npb = [], []
        for x, y in zip(pa, pb):
            if la[x] in gone:
                continue
            npa.append(newpos(la, x))
            npb.append(newpos(lb, y))
        new_steps.append((a, b, npa, npb))
    return new_ranks, new_steps, {t: sorted(c) for t, c in cuts.items()}


def hoist_invariant(n_inputs, steps, variant_inputs):
    """Stable re-ordering of a plan: the steps that do not depend (transitively) on any tensor in `variant_inputs`
    first, the dependent ones after them, tensor ids renumbered accordingly.  Every step is still the same pairwise
    contraction of the same operands - only independent steps swap places - so the values are bit-identical; the
    invariant prefix is what all slices of a network share (qtb_plan_create_sliced runs it once per amplitude).
    Returns (steps', n_invariant_steps)."""
    dep = set(variant_inputs)
    order_inv, order_dep = [], []
    for i, (a, b, _, _) in enumerate(steps):
        if a in dep or b in dep:
            dep.add(n_inputs + i)
            order_dep.append(i)
        else:
            order_inv.append(i)
    order = order_inv + order_dep
    new_id = {t: t for t in range(n_inputs)}
    for pos, i in enumerate(order):
        new_id[n_inputs + i] = n_inputs + pos
    out = [(new_id[steps[i][0]], new_id[steps[i][1]], list(steps[i][2]), list(steps[i][3])) for i in order]
    return out, len(order_inv)


def slice_inputs(inputs, input_ranks, cuts, wires, digits):
    """input tensors of the slice in which wire wires[i] carries digit digits[i]"""
    value = dict(zip(wires, digits))
    out = []
    for t, x in enumerate(inputs):
        if t not in cuts:
            out.append(np.asarray(x, dtype=np.complex128))
            continue
        arr = np.asarray(x, dtype=np.complex128).reshape((4,) * input_ranks[t], order="F")
        idx = [slice(None)] * input_ranks[t]
        for (l, w) in cuts[t]:
            idx[l] = value[w]
        out.append(np.ascontiguousarray(arr[tuple(idx)].reshape(-1, order="F")))
    return out


def all_slices(wires):
    return list(itertools.product(range(4), repeat=len(wires)))


def compile_sliced(engine, input_ranks, steps, wires):
    """One device plan for all slices: sliced legs dropped, slice-invariant steps hoisted into a run-once prefix.
    Returns (plan, cuts, n_invariant_steps)."""
    ranks2, steps2, cuts = slice_plan(input_ranks, steps, wires)
    steps3, n_inv = hoist_invariant(len(ranks2), steps2, cuts.keys())
    return engine.plan(ranks2, steps3, invariant_steps=n_inv), cuts, n_inv


def contract_sliced(engine, input_ranks, steps, inputs, wires, dispatcher=None):
    """Sum over the 4^s slices on the device: one compiled plan, every owned slice staged as an input slot; the steps no
    cut wire reaches run once, the others once per owned slice, partial sums meet in one allreduce."""
    from .dispatch import Dispatcher
    dispatcher = dispatcher or Dispatcher()
    plan, cuts, n_inv = compile_sliced(engine, input_ranks, steps, wires)
    slices = all_slices(wires)
    owned = dispatcher.owned(len(slices))
    for slot, u in enumerate(owned):
        plan.stage_inputs(slot, slice_inputs(inputs, input_ranks, cuts, wires, slices[u]))
    partial = plan.run_slots(range(len(owned))) if owned else 0.0 + 0.0j
    total = dispatcher.map_reduce(dispatcher.world, lambda u: partial)      # one partial per rank, one allreduce
    info = {"slices": len(slices), "owned": len(owned), "units_per_slice": plan.units - plan.prefix_units,
            "units_shared": plan.prefix_units, "invariant_steps": n_inv, "launches_per_slice": plan.launches,
            "peak_rank": plan_cost(input_ranks, steps, frozenset(wires))[1]}
    plan.destroy()
    return total, info
