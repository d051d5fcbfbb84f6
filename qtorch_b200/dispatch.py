"""Term / slice dispatcher: shards INDEPENDENT units of work over the ranks of one job and combines the scalars
with a single sum-allreduce.

Replaces the serial per-edge loop of the reference's QAOA objective (`f_pVal += 0.5*(1-Re<ZZ>)`,
/root/reference/src/maxcut.cpp:171-198) and provides the reduction for index-sliced sub-networks.  The units have no
cross dependence, so there is no data-path collective: unit u goes to rank u % world ("round robin") and
every rank contributes one partial sum.  On GPUs the reduction is `qtb_allreduce_sum` (one ncclAllReduce of complex
scalars over NVLink, issued from the engine stream); the same logic runs over gloo for the CPU test-suite.
"""
import numpy as np


def deal_round_robin(n_units, rank, world):
    """indices of the units rank `rank` owns"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_units, world))


class Dispatcher:
    """evaluate(u) -> complex for every owned unit; reduce partial sums (and, optionally, the per-unit vector)."""

    def __init__(self, rank=0, world=1, allreduce=None):
        self.rank, self.world = rank, world
        self._allreduce = allreduce            # callable(np.ndarray complex128) -> summed array; None = single rank

    @staticmethod
    def for_engine(engine, rank, world):
        """NCCL reduction through the C ABI (engine.comm_init must have been called when world > 1)"""
        return Dispatcher(rank, world, engine.allreduce_sum if world > 1 else None)

    @staticmethod
    def for_torch_distributed():
        """any torch.distributed backend (gloo in the CPU tests)"""
        import torch
        import torch.distributed as dist

        def allreduce(v):
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.complex128).view(np.float64).copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy().view(np.complex128)

        return Dispatcher(dist.get_rank(), dist.get_world_size(), allreduce if dist.get_world_size() > 1 else None)

    def owned(self, n_units):
        return deal_round_robin(n_units, self.rank, self.world)

    def map_reduce(self, n_units, evaluate, transform=lambda v: v, want_vector=False):
        """sum_u transform(evaluate(u)) over all ranks; with want_vector also the full per-unit vector (zeros elsewhere summed)."""
        vec = np.zeros(n_units if want_vector else 0, dtype=np.complex128)
        partial = 0.0 + 0.0j
        for u in self.owned(n_units):
            val = transform(evaluate(u))
            partial += val
            if want_vector:
                vec[u] = val
        payload = np.concatenate([np.array([partial], dtype=np.complex128), vec])
        if self._allreduce is not None:
            payload = self._allreduce(payload)
        return (payload[0], payload[1:]) if want_vector else payload[0]


def qaoa_objective(dispatcher, n_edges, zz_value):
    """F_p = sum_edges 0.5 * (1 - Re<Z_i Z_j>)  (reference maxcut.cpp:196), edges dealt round-robin"""
    return dispatcher.map_reduce(n_edges, zz_value, transform=lambda v: 0.5 * (1.0 - v.real)).real
