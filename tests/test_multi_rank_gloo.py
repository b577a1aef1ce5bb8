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
