"""CPU, world_size 2, gloo: the N>1 host path — shard bounds, per-shard search with id offsets, ONE all-gather of the
packed lists, merge — with the oracle standing in for the two CUDA kernels (K2 per shard, K3 merge).  The merged
result on every rank must equal the oracle's search over the whole DB."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reverso_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from revers_o_b200.sharded import allgather_packed, pack_results, shard_bounds, unpack_results
        rs = np.random.RandomState(0)                       # same DB / queries on every rank
        n, d, nq, k = 5000, 64, 6, 20                      # 40 row blocks -> shards [0,2560) and [2560,5000)
        db = O.round_to_bf16(O._cosine_prepare(rs.randn(n, d).astype(np.float32)))
        db[4000] = db[17]                                  # cross-shard tie
        q = rs.randn(nq, d).astype(np.float32)
        q[0] = db[17]
        lo, hi = shard_bounds(n, world, rank)
        ids = np.full((nq, k), -1, np.int64); sc = np.full((nq, k), -np.inf, np.float32); cnt = np.zeros(nq, np.int32)
        for i, (a, b) in enumerate(O.search_batch(db[lo:hi], q, k, 0.0, db_is_normalized=True)):
            ids[i, : len(a)], sc[i, : len(a)], cnt[i] = a + lo, b, len(a)
        blob = pack_results(torch.from_numpy(ids), torch.from_numpy(sc), torch.from_numpy(cnt))
        gathered = allgather_packed(blob)
        gi, gs, gc = unpack_results(gathered, nq, k)
        mi, ms, mc = O.merge_topk(gi.numpy(), gs.numpy(), gc.numpy(), k)
        ok = True
        for i, (a, b) in enumerate(O.search_batch(db, q, k, 0.0, db_is_normalized=True)):
            ok &= mc[i] == len(a) and np.allclose(ms[i, : len(a)], b, atol=1e-6)
            ok &= set(mi[i, : len(a)].tolist()) == set(a.tolist())
        ok &= mi[0, :2].tolist() == [17, 4000]             # tie across shards -> lower global id first
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_gather_merge():
    world = 2
    port = _free_port()
    with mp.get_context("spawn").Manager() as man:      # no fork() of the multi-threaded pytest process
        ret = man.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
