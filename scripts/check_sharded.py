"""Multi-GPU check (run under torchrun, one rank per GPU): the row-sharded search must equal the
unsharded search bit for bit over every exchange transport —
  collective : local top-k -> one NCCL all-gather -> merge kernel
  peer/symm  : local top-k -> ONE kernel (NVLink stores into every peer + flag + wait + merge),
               buffers mapped with torch symmetric memory
  peer/ipc   : same kernel, buffers mapped with CUDA IPC (irr_peer_export / irr_peer_import)
and a latency comparison of the transports (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_sharded.py [--time]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("IRR_EXCHANGE_TIMEOUT_MS", "20000")   # a check script: trap early, do not hang

import torch
import torch.distributed as dist

import imageretrievalresearch_b200 as irr

TRANSPORTS = [("collective", None), ("peer", "symm"), ("peer", "ipc")]


def make_sharded(full, exchange, mapping, **kw):
    if mapping is not None:
        os.environ["IRR_PEER_MAPPING"] = mapping
    else:
        os.environ.pop("IRR_PEER_MAPPING", None)
    return irr.ShardedGallery.from_full(full, exchange=exchange, **kw)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    usable = []
    for exchange, mapping in TRANSPORTS:
        name = exchange if mapping is None else f"{exchange}/{mapping}"
        try_ok = True
        for (N, D, Q, k, dt) in [(200_003, 1536, 300, 3, torch.bfloat16),
                                 (50_000, 2560, 64, 10, torch.bfloat16),
                                 (10_000, 1536, 64, 3, torch.float32), (5, 64, 7, 3, torch.float32),
                                 (20_000, 256, 33, 150, torch.bfloat16)]:
            gen = torch.Generator(device=dev).manual_seed(1234)          # same data on every rank
            full = torch.randn(N, D, device=dev, generator=gen).to(dt)
            full[N // 2] = full[1]                                      # a cross-shard exact tie
            qs = [torch.randn(Q, D, device=dev, generator=gen).to(dt) for _ in range(5)]
            for q in qs:
                q[0] = full[1].float() * 2
            try:
                sg = make_sharded(full, exchange, mapping, cache_norms=(k == 3))
                # five back-to-back searches without a host sync: exercises both buffer halves and
                # ranks running one call ahead of each other
                gots = [sg.search(q, k) for q in qs]
            except RuntimeError as e:
                if rank == 0:
                    print(f"[SKIP] {name}: {e}", flush=True)
                try_ok = False
                break
            same = True
            # the unsharded search with the SAME norm source as the shards (cached norms come from
            # the row-norm kernel, uncached ones from the top-k kernel's own warps: identical
            # ranking, last-bit differences in the values)
            whole = irr.Gallery(full, cache_norms=(k == 3))
            for q, got in zip(qs, gots):
                want = whole.search(q, k)
                same &= torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"[{'PASS' if flag.item() else 'FAIL'}] {name} x{world} == unsharded: N={N} D={D} "
                      f"Q={Q} k={k} {dt} transport={sg.transport}; tie row -> "
                      f"{gots[0].indices[0, :2].tolist()}", flush=True)
            ok &= bool(flag.item())
            if exchange == "peer" and k <= 16:
                # lagged stream: same results, one call late; ranks deliberately out of step
                outs = []
                for j, q in enumerate(qs):
                    if rank == j % world:
                        torch.cuda._sleep(20_000_000)      # ~10 ms: this rank falls behind
                    r = sg.search_lagged(q, k)
                    if r is not None:
                        outs.append(r)
                outs.append(sg.flush())
                same = all(torch.equal(o.indices, g.indices) and torch.equal(o.values, g.values)
                           for o, g in zip(outs, gots)) and len(outs) == len(gots)
                flag = torch.tensor([1 if same else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print(f"[{'PASS' if flag.item() else 'FAIL'}] {name} x{world} lagged stream == plain "
                          f"search: N={N} Q={Q} k={k}", flush=True)
                ok &= bool(flag.item())
            sg.close()
        if try_ok:
            usable.append((exchange, mapping))

    if "--time" in sys.argv:
        N, D, k = 1_000_000, 1536, 3
        lo, hi = irr.shard_bounds(N, world, rank)
        shard = torch.randn(hi - lo, D, device=dev).to(torch.bfloat16)

        def timed(fn, iters):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item()

        for Q in (1, 64, 4096):
            q = torch.randn(Q, D, device=dev).to(torch.bfloat16)
            arms = {}
            for exchange, mapping in usable:
                if mapping == "ipc" and ("peer", "symm") in usable:
                    continue                      # same kernel; one mapping is enough for timing
                if mapping is not None:
                    os.environ["IRR_PEER_MAPPING"] = mapping
                sg = irr.ShardedGallery(shard, N, exchange=exchange)
                arms[exchange] = (sg, (lambda sg=sg: sg.search(q, k)))
                if exchange == "peer":
                    cap = sg.capture(Q, k)
                    cap(q)
                    arms["peer+graph"] = (None, (lambda cap=cap: cap()))
            iters = 200 if Q <= 64 else 30
            best = {name: [] for name in arms}
            for name, (_, fn) in arms.items():
                for _ in range(10):
                    fn()
            for _ in range(3):                    # interleaved rounds: no arm always runs "first"
                for name, (_, fn) in arms.items():
                    best[name].append(timed(fn, iters))
            if rank == 0:
                for name, ts in best.items():
                    print(json.dumps({"Q": Q, "world": world, "arm": name,
                                      "ms_per_search_min": round(min(ts), 4),
                                      "ms_per_search_all": [round(t, 4) for t in ts],
                                      "queries_per_s": round(Q / min(ts) * 1e3, 1)}), flush=True)
            dist.barrier()
            for sg, _ in arms.values():
                if sg is not None:
                    sg.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
