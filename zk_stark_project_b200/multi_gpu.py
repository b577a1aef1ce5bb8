"""Multi-GPU host logic: one process per GPU, launched by torchrun; torch.distributed is plumbing only.

Independent proofs (BASELINE.json configs 2-4) shard across ranks with no data-path collective: rank r proves the
proofs `shard_proofs(...)` assigns to it on its own device and the results are gathered for the caller."""
import hashlib


def shard_proofs(num_proofs, world_size, rank):
    """Round-robin assignment of proof indices to ranks (proof i -> rank i % world_size)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, num_proofs, world_size))


def gather_proofs(local, num_proofs, dist=None):
    """local: {proof index: proof bytes} of this rank -> list of all proofs in index order on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local[i] for i in range(num_proofs)]
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, local)
    merged = {}
    for part in gathered:
        merged.update(part)
    missing = [i for i in range(num_proofs) if i not in merged]
    if missing:
        raise RuntimeError(f"proofs {missing} were not produced by any rank")
    return [merged[i] for i in range(num_proofs)]


def max_over_ranks(value, dist=None, device=None):
    """Elapsed time of a multi-rank step = the slowest rank's (bench.py contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def prove_batch(provers_and_traces, ctx, rank=0, world_size=1, dist=None):
    """Config 4: a batch of independent proofs, proof i on rank i % world_size; returns all proofs on every rank."""
    import numpy as np
    mine = {}
    for i in shard_proofs(len(provers_and_traces), world_size, rank):
        prover, trace = provers_and_traces[i]
        data = np.ascontiguousarray(trace.data)
        proof, _ = ctx.prove_host(prover.describe(trace), data.ctypes.data)
        mine[i] = proof
    return gather_proofs(mine, len(provers_and_traces), dist)


def digest(proof):
    return hashlib.sha256(proof).hexdigest()
