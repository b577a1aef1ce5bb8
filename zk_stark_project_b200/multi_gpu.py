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


def prove_batch(provers_and_traces, ctx, rank=0, world_size=1, dist=None, lanes=None):
    """Config 4: a batch of independent proofs, proof i on rank i % world_size; returns all proofs on every rank.
    `lanes`: extra zkb contexts on this rank's device — this rank's share then goes through zkb_prove_batch, which keeps
    len(lanes) + 1 proofs in flight (host threads inside the library)."""
    import numpy as np
    from . import lib as L
    idx = list(shard_proofs(len(provers_and_traces), world_size, rank))
    airs, datas = [], []
    for i in idx:
        prover, trace = provers_and_traces[i]
        airs.append(prover.describe(trace))
        datas.append(np.ascontiguousarray(trace.data))
    if lanes:
        proofs = L.prove_batch([ctx] + list(lanes), airs, [d.ctypes.data for d in datas])
    else:
        proofs = [ctx.prove_host(a, d.ctypes.data)[0] for a, d in zip(airs, datas)]
    return gather_proofs(dict(zip(idx, proofs)), len(provers_and_traces), dist)


def digest(proof):
    return hashlib.sha256(proof).hexdigest()


# ---- column-sharded single proof (BASELINE.json configs[4]): index logic shared by the C++ driver and its tests ---------
def column_shard(width, world_size, rank):
    """Columns [r*w/G, (r+1)*w/G) of rank r; G must be a power of two dividing w (zkb_mg_prove)."""
    if width % world_size or world_size & (world_size - 1):
        raise ValueError("world_size must be a power of two that divides the trace width")
    wl = width // world_size
    return range(rank * wl, (rank + 1) * wl)


def row_shard(lde_size, world_size, rank):
    """LDE rows (= Merkle leaves) [q*N/G, (q+1)*N/G) hashed by rank q after the all-to-all."""
    nl = lde_size // world_size
    return range(rank * nl, (rank + 1) * nl)


def node_owner(heap_index, lde_size, world_size):
    """Owner of Merkle heap node `heap_index` (root = 1, leaves = N..2N-1): -1 for the replicated cap (top log2 G levels),
    else (rank, index in that rank's local subtree heap)."""
    log_g = world_size.bit_length() - 1
    d = heap_index.bit_length() - 1
    if d <= log_g:
        return -1, heap_index
    off = heap_index - (1 << d)
    return off >> (d - log_g), (1 << (d - log_g)) + (off & ((1 << (d - log_g)) - 1))


def finish_cap(subtree_roots, merge):
    """Top log2 G levels from the all-gathered subtree roots; returns the heap (cap[1] = root)."""
    g = len(subtree_roots)
    cap = [None] * (2 * g)
    cap[g:] = list(subtree_roots)
    for i in range(g - 1, 0, -1):
        cap[i] = merge(cap[2 * i], cap[2 * i + 1])
    return cap


def init_sharded(ctx, rank, world_size, dist):
    """Create the library's NCCL communicator: rank 0 draws the id, torch.distributed broadcasts it (control plane only)."""
    from . import lib
    ids = [lib.mg_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.mg_init(rank, world_size, ids[0])
