"""Multi-GPU parity check (launch with torchrun, one rank per GPU): row-sharded search over NCCL == the CPU oracle's
search over the whole DB, and == the single-GPU search of the unsharded DB.  Prints one OK/FAIL line on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/check_sharded_nccl.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import reverso_oracle as O          # checker only
from revers_o_b200 import ops, synth
from revers_o_b200.sharded import ShardedIndex, selfjoin_blocks, shard_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    msgs = []
    push_note = []
    for (n, d, nq, k, thr) in [(200_000, 1024, 64, 100, None), (150_001, 1280, 300, 50, 0.6), (60_000, 256, 3, 10, None)]:
        q = synth.make_queries(nq, d, seed=7, device=dev)
        full = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)      # same DB on every rank (same seed)
        lo, hi = shard_bounds(n, world, rank)
        local_db = full[lo // 128: (hi + 127) // 128].contiguous()
        idx = ShardedIndex(local_db, hi - lo, d, lo)
        ids, sc, cnt = idx.search_exact(q, k, thr)
        ids, sc, cnt = ids.clone(), sc.clone(), cnt.clone()
        # the same search through the peer-memory exchange (NVLink stores fused into K2's last kernel) must agree bit for bit,
        # repeatedly (both slot parities, reuse of the flags)
        pushed = idx.enable_peer_exchange(nq_max=max(nq, 8), k_max=k)
        if pushed and nq > 4:
            # the separate-merge protocol (push + rvo_merge_topk_exchange) must agree with the fused one used by default
            os.environ["RVO_FUSED"] = "0"
            ui, us, uc = (t.clone() for t in idx.search(q, k, thr))
            os.environ["RVO_FUSED"] = "1"
            fi_, fs_, fc_ = (t.clone() for t in idx.search(q, k, thr))
            torch.cuda.synchronize()
            if not (torch.equal(ui, fi_) and torch.equal(us, fs_) and torch.equal(uc, fc_)):
                ok = False
                msgs.append("fused exchange != separate merge")
        if pushed:
            for _ in range(5):
                pi, ps, pc = idx.search(q, k, thr)
                torch.cuda.synchronize()
                bad = (cnt < 0)
                same_push = bool(torch.equal(pc[~bad], cnt[~bad]) and torch.equal(pi[~bad], ids[~bad]) and torch.equal(ps[~bad], sc[~bad]))
                if not same_push:
                    break
            # pipelined serving (submit / collect, scan of batch i+1 enqueued before the merge of batch i): nine DIFFERENT query
            # batches through the four slot sets — a stale or overwritten slot would surface as a wrong list
            if same_push and nq > 4:
                variants = [q.roll(j, dims=0).contiguous() for j in range(9)]
                want = [tuple(t.clone() for t in idx.search(v, k, thr)) for v in variants]
                got, prev = [], None
                for v in variants:
                    t = idx.submit(v, k, thr)
                    if prev is not None:
                        got.append(tuple(t_.clone() for t_ in idx.collect(prev)))
                    prev = t
                got.append(tuple(t_.clone() for t_ in idx.collect(prev)))
                torch.cuda.synchronize()
                same_push = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) for a, b in zip(want, got))
            idx.disable_peer_exchange()
        else:
            same_push = True
        push_note.append(f"push={'on' if pushed else 'unavailable'}:{'ok' if same_push else 'FAIL'}")
        ok &= same_push
        fi, fs, fc = ops.search_topk_exact(full, n, d, q, k, thr)
        torch.cuda.synchronize()
        same = bool(torch.equal(ids, fi) and torch.equal(cnt, fc) and torch.allclose(sc, fs, atol=1e-6))
        if rank == 0:
            sel = list(range(0, nq, max(1, nq // 8)))
            ref = O.search_batch(ops.untile_rows(full, n, d).float().cpu().numpy(), q[sel].cpu().numpy(), k, thr,
                                 db_is_normalized=True)
            for j, (rid, rsc) in enumerate(ref):
                c = int(cnt[sel[j]])
                good = c == len(rid) and (c == 0 or (np.max(np.abs(sc[sel[j], :c].cpu().numpy() - rsc)) < 1e-3 and
                                                     set(ids[sel[j], :c].tolist()) == set(rid.tolist())))
                same &= bool(good)
        t = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok &= bool(t.item())
        msgs.append(f"n{n}d{d}q{nq}k{k}thr{thr}:{'ok' if t.item() else 'FAIL'}")

    # self-join: DB replicated, query blocks dealt across ranks, pair counts add up to the single-GPU join
    n, d = 40_000, 1024
    db = synth.make_db(n, d, None, seed=5, device=dev)
    rows = ops.untile_rows(db, n, d)
    dup = torch.arange(0, 2000, device=dev)
    rows2 = rows.clone()
    rows2[20_000 + dup] = rows[dup]                                            # 2000 exact duplicates
    db = ops.tile_rows(rows2)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    for (a, b) in selfjoin_blocks(n, world, rank):
        pairs, scores, count, over = ops.selfjoin_threshold(db, n, d, 0.95, row_lo=a, row_hi=b)
        total += int(count.item())
    dist.all_reduce(total)
    ok &= int(total.item()) == 2000
    msgs.append(f"selfjoin pairs {int(total.item())}/2000")
    if rank == 0:
        print(("SHARDED_NCCL_OK " if ok else "SHARDED_NCCL_FAIL ") + f"world={world} " + " ".join(msgs) + " " + " ".join(push_note),
              flush=True)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
