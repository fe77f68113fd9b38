"""world_size-2 (and 3) runs of the sharded search's host logic over gloo on the CPU: shard bounds,
index offsets, the single all-gather exchange and the merge order.  The per-shard top-k and the
merge are done by the ORACLE here (the CUDA kernels need a GPU; tests/test_gpu_parity.py runs the
same flow with the kernels, emulating the ranks on one device)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_path as ref
from oracle import synthetic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from imageretrievalresearch_b200 import exchange_candidates, shard_bounds
        q, gal = synthetic.tied_gallery(N, 40, 12, seed=13)
        lo, hi = shard_bounds(N, world, rank)
        kk = min(k, hi - lo)
        # per-shard top-k from the same score matrix the unsharded oracle ranks (a BLAS call on a
        # different shape may round duplicate rows differently, which is not what is under test)
        s = ref.cos_scores(q, gal).float()
        dist.broadcast(s, src=0)   # MKL may round differently per process (alignment-dependent paths)
        sv, si = torch.sort(s[:, lo:hi], dim=1, descending=True, stable=True)
        v, i = sv[:, :kk], si[:, :kk]
        pad = k - kk
        v = torch.cat([v.float(), torch.full((12, pad), -float("inf"))], 1)
        i = torch.cat([i + lo, torch.full((12, pad), -1, dtype=torch.int64)], 1)
        cv, ci = exchange_candidates(v, i)          # the product's exchange step, over gloo
        assert cv.shape == (world, 12, k) and ci.shape == (world, 12, k)
        mv, mi = ref.merge_candidates(cv, ci, k)
        want_i = torch.sort(s, dim=1, descending=True, stable=True)[1][:, :k]
        ok = torch.equal(mi, want_i) and torch.equal(mv, s.gather(1, want_i))
        # every rank must hold the same merged answer
        if not ok:
            print(f"rank {rank}: idx equal {torch.equal(mi, want_i)}; first rows {mi[0].tolist()} vs "
                  f"{want_i[0].tolist()}", flush=True)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,k", [(2, 501, 3), (2, 5, 3), (3, 1000, 10)])
def test_sharded_search_host_logic_gloo(world, N, k):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == 1
