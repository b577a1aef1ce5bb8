"""CPU, world_size 2 over gloo: the N>1 host logic — proof sharding, gathering and max-over-ranks timing.
The proving itself is stood in for by the oracle here (the GPU box runs the same logic with libzkb200.so)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from zk_stark_project_b200 import multi_gpu as M


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as O
        from tests import common as T
        O.set_threads(2)
        import zk_stark_project_b200 as Z
        provers = [Z.MimcProver(T.options(blowup=8, grinding=4), [100 * i + j + 1 for j in range(1 + i % 3)], 64) for i in range(5)]
        mine = {}
        for i in M.shard_proofs(len(provers), world, rank):
            tr = provers[i].build_trace()
            mine[i] = O.prove(provers[i].describe(tr), tr.to_bytes())[0]
        proofs = M.gather_proofs(mine, len(provers), dist)
        for p, proof in zip(provers, proofs):  # every rank ends up with every proof, in order, and they verify
            O.verify(p.describe(p.build_trace()), proof)
        slowest = M.max_over_ranks(10.0 + rank, dist)
        q.put((rank, sorted(mine), [M.digest(p) for p in proofs], slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_proof_sharding():
    assert M.shard_proofs(5, 2, 0) == [0, 2, 4] and M.shard_proofs(5, 2, 1) == [1, 3]
    assert M.shard_proofs(3, 8, 5) == [] and M.shard_proofs(256, 8, 7)[:2] == [7, 15]
    with pytest.raises(ValueError):
        M.shard_proofs(4, 2, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    assert res[0][2] == res[1][2] and len(set(res[0][2])) == 5  # same gathered proofs on both ranks
    assert res[0][3] == res[1][3] == 11.0                        # max over ranks


def test_single_process_paths():
    assert M.gather_proofs({0: b"a", 1: b"b"}, 2) == [b"a", b"b"]
    assert M.max_over_ranks(3.5) == 3.5


def _sharded_worker(rank, world, port, q):
    """CPU emulation of zkb_mg_prove's trace commitment: column-local LDE -> exchange -> row-shard hashing -> subtree roots
    -> all-gather -> cap; plus the partial-sum constraint evaluation.  Arithmetic comes from the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from oracle import pyoracle as O
        import zk_stark_project_b200 as Z
        O.set_threads(1)
        w, n, beta = 8, 64, 8
        N = n * beta
        rc = Z.get_round_constants()
        trace = np.frombuffer(O.mimc_trace([j + 1 for j in range(w)], n, rc), dtype=np.uint64).reshape(w, n, 2)
        full_root, full_lde, _ = O.trace_commit(trace.tobytes(), n, w, beta, want_lde=True)
        cols = M.column_shard(w, world, rank)
        # column-local interpolation + LDE (what K1/K2 do on this rank's columns)
        _, local_lde, _ = O.trace_commit(np.ascontiguousarray(trace[cols.start:cols.stop]).tobytes(), n, len(cols), beta, want_lde=True)
        local = np.frombuffer(local_lde, dtype=np.uint8).reshape(N, len(cols) * 16)
        # "all-to-all": chunk q of my column shard goes to rank q
        send = {q: local[M.row_shard(N, world, q).start:M.row_shard(N, world, q).stop].tobytes() for q in range(world)}
        gathered = [None] * world
        dist.all_gather_object(gathered, send)
        rows = M.row_shard(N, world, rank)
        mine = np.concatenate([np.frombuffer(gathered[src][rank], dtype=np.uint8).reshape(len(rows), -1) for src in range(world)], axis=1)
        ref = np.frombuffer(full_lde, dtype=np.uint8).reshape(N, w * 16)[rows.start:rows.stop]
        assert np.array_equal(mine, ref), "row shard after the exchange differs from the full LDE"
        leaves = [O.blake3(mine[i].tobytes()) for i in range(len(rows))]
        sub_root = O.merkle_root(leaves)
        roots = [None] * world
        dist.all_gather_object(roots, sub_root)
        cap = M.finish_cap(roots, lambda a, b: O.blake3(a + b))
        assert cap[1] == full_root, "cap over the subtree roots must equal the single-prover trace root"
        q.put((rank, cap[1].hex()))
    finally:
        dist.destroy_process_group()


def test_two_rank_column_sharded_commitment():
    assert list(M.column_shard(64, 8, 3)) == list(range(24, 32)) and list(M.row_shard(1 << 10, 4, 1))[:2] == [256, 257]
    assert list(M.column_shard(240, 8, 7)) == list(range(210, 240))  # the training width splits 8 x 30
    with pytest.raises(ValueError):
        M.column_shard(240, 7, 0)
    with pytest.raises(ValueError):
        M.column_shard(64, 6, 0)
    # heap ownership: N = 16 leaves, G = 4: depth <= 2 replicated, below that owned by leaf range
    assert M.node_owner(1, 16, 4) == (-1, 1) and M.node_owner(7, 16, 4) == (-1, 7)
    assert M.node_owner(8, 16, 4) == (0, 2) and M.node_owner(15, 16, 4) == (3, 3) and M.node_owner(16 + 5, 16, 4) == (1, 4 + 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]
