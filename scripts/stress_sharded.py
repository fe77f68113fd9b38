"""Randomised multi-GPU stress (under torchrun): random (N, D, Q, k, dtype) galleries, row-sharded
over the ranks through the default (peer-memory) exchange, several back-to-back searches without a
host sync — each result must equal the unsharded search of the full gallery bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29571 scripts/stress_sharded.py [seconds] [seed]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("IRR_EXCHANGE_TIMEOUT_MS", "20000")

import torch
import torch.distributed as dist

import imageretrievalresearch_b200 as irr


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    rnd = torch.Generator().manual_seed(seed)          # the same sequence on every rank

    def ri(lo, hi):
        return int(torch.randint(lo, hi + 1, (1,), generator=rnd).item())

    t0 = time.time()
    cases, searches = 0, 0
    stop = torch.zeros(1, device=dev)
    while True:
        stop.fill_(1.0 if time.time() - t0 > secs else 0.0)
        dist.all_reduce(stop, op=dist.ReduceOp.MAX)     # all ranks leave together
        if stop.item() > 0:
            break
        dt = (torch.bfloat16, torch.float32)[ri(0, 3) == 0]
        N = (ri(1, 40), ri(100, 5000), ri(5000, 120_000))[ri(0, 2)]
        D = 8 * ri(1, 64) if ri(0, 1) else (64, 1536)[ri(0, 1)]
        Q = (ri(1, 64), ri(1, 300), ri(300, 1100))[ri(0, 2)]
        k = min((3, ri(1, 16), ri(17, 200))[ri(0, 2)], N)
        if dt == torch.float32:
            N, Q = min(N, 20_000), min(Q, 200)
        cached = bool(ri(0, 1))
        gen = torch.Generator(device=dev).manual_seed(seed * 7919 + cases)   # same data on every rank
        full = torch.randn(N, D, device=dev, generator=gen).to(dt)
        if N >= 4:
            full[N // 2] = full[1]                      # an exact tie across the shard boundary
        qs = [torch.randn(Q, D, device=dev, generator=gen).to(dt) for _ in range(3)]
        sg = irr.ShardedGallery.from_full(full, cache_norms=cached)
        gots = [sg.search(q, k) for q in qs]            # back to back, no host sync in between
        whole = irr.Gallery(full, cache_norms=cached)
        same = True
        for q, got in zip(qs, gots):
            want = whole.search(q, k)
            same &= torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not flag.item():
            if rank == 0:
                print("MISMATCH", dict(N=N, D=D, Q=Q, k=k, dtype=str(dt), cached=cached, world=world), flush=True)
            dist.destroy_process_group()
            sys.exit(1)
        cases += 1
        searches += len(qs)
        del sg
    if rank == 0:
        print(f"ok: {cases} random sharded cases ({searches} searches) on {world} GPUs in {time.time() - t0:.0f} s", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
