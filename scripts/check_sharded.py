"""Multi-GPU check (run under torchrun, one rank per GPU): the row-sharded search
(local top-k -> one NCCL all-gather -> merge kernel) must equal the unsharded search bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_sharded.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import imageretrievalresearch_b200 as irr


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for (N, D, Q, k, dt) in [(200_003, 1536, 300, 3, torch.bfloat16), (50_000, 2560, 64, 10, torch.bfloat16),
                             (10_000, 1536, 64, 3, torch.float32), (5, 64, 7, 3, torch.float32)]:
        gen = torch.Generator(device=dev).manual_seed(1234)          # same data on every rank
        full = torch.randn(N, D, device=dev, generator=gen).to(dt)
        full[N // 2] = full[1]                                      # a cross-shard exact tie
        q = torch.randn(Q, D, device=dev, generator=gen).to(dt)
        q[0] = full[1].float() * 2
        sg = irr.ShardedGallery.from_full(full, cache_norms=(k == 3))
        got = sg.search(q, k)
        want = irr.cosine_topk(q, full, k)
        same = torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[{'PASS' if flag.item() else 'FAIL'}] sharded x{world} == unsharded: N={N} D={D} Q={Q} "
                  f"k={k} {dt}; tie row -> {got.indices[0, :2].tolist()}", flush=True)
        ok &= bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
