"""Launch-shape sweep of the fused loss kernel (warps per row x ring depth), one subprocess per
combination because the library reads its measurement knobs once per process.

    python scripts/loss_sweep.py            # all combinations, one JSON line each
    python scripts/loss_sweep.py one        # measure the current environment's shape only
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one():
    import torch
    import imageretrievalresearch_b200 as irr
    from bench import graphed_us, measured_peaks
    peaks = measured_peaks()
    B, D = 4096, 1536
    out = {"gw": os.environ.get("IRR_LOSS_GW", "auto"), "stages": os.environ.get("IRR_LOSS_STAGES", "auto")}
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        sets = [[torch.randn(B, D, device="cuda").to(dt) for _ in range(3)] for _ in range(6)]
        us = min(graphed_us(lambda i: irr.triplet_losses_fwd_bwd(*sets[i % 6], 0.3), 24) for _ in range(3))
        by = 6 * B * D * sets[0][0].element_size()
        out[name + "_us"] = round(us, 2)
        out[name + "_hbm_frac"] = round(by / us / 1e3 / peaks["hbm_gbs"], 3)
        del sets
    B = 65536
    q, p, n = [torch.randn(B, D, device="cuda") for _ in range(3)]
    us = graphed_us(lambda i: irr.triplet_losses_fwd_bwd(q, p, n, 0.3), 4)
    out["f32_64k_hbm_frac"] = round(6 * B * D * 4 / us / 1e3 / peaks["hbm_gbs"], 3)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        combos = [(None, None)] + [(gw, st) for gw in (1, 2, 4) for st in (2, 3, 4)]
        for gw, st in combos:
            env = dict(os.environ)
            if gw is not None:
                env["IRR_LOSS_GW"], env["IRR_LOSS_STAGES"] = str(gw), str(st)
            r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"gw": gw, "stages": st, "error": r.stderr[-400:]}) + "\n")
            sys.stdout.flush()
