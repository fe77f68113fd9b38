"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF in the authoring container.

    python oracle/gen_golden.py            # needs /root/reference (read-only), CPU only

The reference's only importable first-party code on this path is ``ContrastiveLoss``
(utils/contrastive_loss.py); it is imported here by file path, unmodified.  Everything else on the
path is a torch call the reference scripts make inline (CosineSimilarity(dim=1, eps=1e-6),
torch.topk, torch.nn.CosineEmbeddingLoss(margin)) — those calls are made here with the arguments
and label tensors the reference uses (train/train_efficient_cos_con_ce_loss.py:89,97-100,158-159,
230-237,270-281,377-382).  The vectors are small so that they can be committed; the GPU box has no
/root/reference, so tests only ever read the .npz files.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import synthetic  # noqa: E402

REF = Path(os.environ.get("IRR_REFERENCE", "/root/reference"))
OUT = ROOT / "tests" / "golden"


def load_reference_contrastive():
    spec = importlib.util.spec_from_file_location("ref_contrastive_loss", REF / "utils" / "contrastive_loss.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ContrastiveLoss


def golden_losses(ContrastiveLoss):
    out = {}
    labels = {"con_pos": torch.tensor(1.).unsqueeze(0), "con_neg": torch.tensor(0.).unsqueeze(0),
              "cos_pos": torch.tensor(1.).unsqueeze(0), "cos_neg": torch.tensor(-1.).unsqueeze(0)}
    for tag, (B, D, scaled) in {"unit": (48, 96, False), "scaled": (33, 128, True)}.items():
        q, p, n = synthetic.triplets(B, D, seed=11, scaled=scaled)
        out[f"{tag}_q"], out[f"{tag}_p"], out[f"{tag}_n"] = q.numpy(), p.numpy(), n.numpy()
        for margin in (0.2, 0.3, 0.5):
            qr, pr, nr = [t.clone().requires_grad_(True) for t in (q, p, n)]
            cos_loss = torch.nn.CosineEmbeddingLoss(margin=margin)      # reference :158
            con_loss = ContrastiveLoss(margin=margin)                   # reference :159
            l = torch.stack([cos_loss(qr, pr, labels["cos_pos"]), cos_loss(qr, nr, labels["cos_neg"]),
                             con_loss(qr, pr, labels["con_pos"]), con_loss(qr, nr, labels["con_neg"])])
            (l[0] + l[1] + l[2] + l[3]).backward()                      # loss_cos + loss_con, :232,237,245
            key = f"{tag}_m{margin}"
            out[key + "_losses"] = l.detach().numpy()
            out[key + "_dq"], out[key + "_dp"], out[key + "_dn"] = qr.grad.numpy(), pr.grad.numpy(), nr.grad.numpy()
            out[key + "_con_sum"] = np.array([con_loss(q, p, 1., mean=False).item(),
                                              con_loss(q, n, 0., mean=False).item()], dtype=np.float32)
        cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)                # reference :89
        sims = torch.stack([cos(q[i].unsqueeze(0), p[i].unsqueeze(0))[0] for i in range(B)])   # :377
        unsims = torch.stack([cos(q[i].unsqueeze(0), n[i].unsqueeze(0))[0] for i in range(B)])  # :381
        out[f"{tag}_cos_sims"], out[f"{tag}_cos_unsims"] = sims.numpy(), unsims.numpy()
    # the docstring example shape of utils/contrastive_loss.py:62-65: rand(3,4,4) -> scalar
    g = torch.Generator().manual_seed(5)
    a, b = torch.rand(3, 4, 4, generator=g), torch.rand(3, 4, 4, generator=g)
    out["doc_a"], out["doc_b"] = a.numpy(), b.numpy()
    out["doc_loss"] = np.array(ContrastiveLoss(0.5)(a, b, 1).item(), dtype=np.float32)
    return out


def golden_autocast(ContrastiveLoss):
    """fp16 embeddings under precision=16 (train/train_efficient_cos_con_ce_loss.py:465): the
    reference's `fm2 - fm1` (utils/contrastive_loss.py:56) is an fp16 subtraction; pow / sum and
    cosine_embedding_loss are widened to fp32 by autocast.  CPU autocast has a
    different op list, so the policy is applied by hand AROUND the unmodified reference module: it
    is given fm1 = 0 and fm2 = the fp16-rounded difference (fm2 - fm1 is then exactly that
    difference), i.e. the module evaluates precisely what it evaluates under CUDA autocast.
    Tight positives (p = q + 0.002 randn) make the fp16 rounding of p - q matter (> 1e-5)."""
    out = {}
    B, D = 40, 96
    q, p, n = synthetic.triplets(B, D, seed=17)
    p = synthetic.unit(q + 0.002 * torch.randn(B, D, generator=torch.Generator().manual_seed(18)))
    q16, p16, n16 = q.half(), p.half(), n.half()
    out["ac_q"], out["ac_p"], out["ac_n"] = q16.numpy(), p16.numpy(), n16.numpy()
    one, minus = torch.tensor(1.).unsqueeze(0), torch.tensor(-1.).unsqueeze(0)
    zeros = torch.zeros(B, D)
    for margin in (0.2, 0.3, 0.5):
        qr, pr, nr = [t.clone().requires_grad_(True) for t in (q16, p16, n16)]
        cos_loss = torch.nn.CosineEmbeddingLoss(margin=margin)
        con_loss = ContrastiveLoss(margin=margin)
        l = torch.stack([cos_loss(qr.float(), pr.float(), one), cos_loss(qr.float(), nr.float(), minus),
                         con_loss(zeros, (pr - qr).float(), torch.tensor(1.).unsqueeze(0)),
                         con_loss(zeros, (nr - qr).float(), torch.tensor(0.).unsqueeze(0))])
        # GradScaler-style scale so that the fp16 gradients do not underflow (Lightning does this)
        (1024.0 * l.sum()).backward()
        key = f"ac_m{margin}"
        out[key + "_losses"] = l.detach().numpy()
        out[key + "_dq"], out[key + "_dp"], out[key + "_dn"] = qr.grad.numpy(), pr.grad.numpy(), nr.grad.numpy()
        # the same module on the widened inputs: what a "more exact" implementation would return
        out[key + "_con_pos_widened"] = np.float32(con_loss(q16.float(), p16.float(), 1.).item())
    return out


def golden_retrieval():
    out = {}
    cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)
    # planted neighbours: index parity is meaningful (gaps >> tolerance)
    q, g, pos = synthetic.planted_gallery(N=700, D=96, Q=24, k=3, seed=21)
    clss_g = torch.arange(700) % 9
    clss_q = clss_g[pos[:, 0]].clone()
    clss_q[::5] = (clss_q[::5] + 1) % 9  # some misses
    vals, inds, top1, top3 = [], [], 0, 0
    for idx in range(q.shape[0]):                                        # reference :270-281
        sim = cos(q[idx].unsqueeze(0), g)
        v, i = torch.topk(sim, k=3)
        vals.append(v); inds.append(i)
        if clss_q[idx] == clss_g[i[0]] or clss_q[idx] == clss_g[i[1]] or clss_q[idx] == clss_g[i[2]]:
            top3 += 1
        if clss_q[idx] in clss_g[i[0]]:
            top1 += 1
    out.update(planted_q=q.numpy(), planted_g=g.numpy(), planted_pos=pos.numpy(),
               planted_vals=torch.stack(vals).numpy(), planted_inds=torch.stack(inds).numpy(),
               planted_clss_q=clss_q.numpy(), planted_clss_g=clss_g.numpy(),
               planted_top1=np.int64(top1), planted_top3=np.int64(top3))
    # training-step flavour: the gallery is the batch of positives (:385), B=16
    tq, tp, _ = synthetic.triplets(16, 64, seed=31)
    clss = torch.arange(16) % 5
    top1 = top3 = 0
    v10 = []
    for idx in range(16):
        sim = cos(tq[idx].unsqueeze(0), tp)
        v, i = torch.topk(sim, k=3)
        v10.append(v)
        if clss[idx] == clss[i[0]] or clss[idx] == clss[i[1]] or clss[idx] == clss[i[2]]:
            top3 += 1
        if clss[idx] in clss[i[0]]:
            top1 += 1
    out.update(batch_q=tq.numpy(), batch_p=tp.numpy(), batch_clss=clss.numpy(),
               batch_vals=torch.stack(v10).numpy(), batch_top1=np.int64(top1), batch_top3=np.int64(top3))
    # notebook flavour (ipynb:231-251): top-150 + first 3 distinct classes, gallery = all positives
    gq, gp, _ = synthetic.triplets(420, 64, seed=51)
    cls = torch.arange(420) // 70                      # 6 classes x 70 rows, like Sketchy's ~70 / class
    centers = torch.randn(6, 64, generator=torch.Generator().manual_seed(52))
    gq = gq + 0.22 * synthetic.unit(centers)[cls]
    gp = gp * 0 + synthetic.unit(torch.randn(420, 64, generator=torch.Generator().manual_seed(54))) + 0.22 * synthetic.unit(centers)[cls] + 0.05 * torch.randn(420, 64, generator=torch.Generator().manual_seed(53))
    nb_top1 = nb_top3 = 0
    nb_r, nb_i, nb_v, nb_vals150 = [], [], [], []
    for idx, (gt_reg, fm) in enumerate(zip(cls, gq)):
        vals, inds = torch.topk(cos(fm, gp), k=150)
        nb_vals150.append(vals)
        classes = [int(cls[int(ind)]) for ind in inds]
        top_i, top_v, top_r = [], [], []
        for num, (i, v, r) in enumerate(zip(inds, vals, classes)):
            if r not in top_r:
                top_r.append(r); top_v.append(float(v)); top_i.append(int(i))
            if len(top_r) == 3:
                break
        nb_top3 += 1 if int(gt_reg) in top_r else 0
        nb_top1 += 1 if int(gt_reg) == top_r[0] else 0
        pad = 3 - len(top_r)
        nb_r.append(top_r + [-1] * pad); nb_i.append(top_i + [-1] * pad); nb_v.append(top_v + [float("-inf")] * pad)
    out.update(nb_q=gq.numpy(), nb_p=gp.numpy(), nb_cls=cls.numpy(), nb_top1=np.int64(nb_top1),
               nb_top3=np.int64(nb_top3), nb_r=np.array(nb_r, dtype=np.int64),
               nb_i=np.array(nb_i, dtype=np.int64), nb_v=np.array(nb_v, dtype=np.float32),
               nb_vals150=torch.stack(nb_vals150).numpy())
    # iid gallery, k=10: values only (index order near ties is torch.topk's business)
    q2, g2 = synthetic.iid_gallery(N=1000, D=72, Q=6, seed=41)
    v = torch.stack([torch.topk(cos(q2[i].unsqueeze(0), g2), k=10)[0] for i in range(6)])
    out.update(iid_q=q2.numpy(), iid_g=g2.numpy(), iid_vals10=v.numpy())
    return out


def golden_producer_consumer():
    """get_fm (:103-122) and loss_ce (:160,240-242) exactly as the reference calls them."""
    out = {}
    g = torch.Generator().manual_seed(61)
    fm = torch.randn(5, 37, 7, 7, generator=g).requires_grad_(True)
    pool = torch.nn.AvgPool2d((fm.shape[2], fm.shape[3]))
    emb = torch.reshape(pool(fm), (-1, fm.shape[1]))
    up = torch.randn(5, 37, generator=g)
    (emb * up).sum().backward()
    out.update(pool_fm=fm.detach().numpy(), pool_out=emb.detach().numpy(), pool_up=up.numpy(),
               pool_grad=fm.grad.numpy())
    la = torch.randn(19, 125, generator=g).mul(3).requires_grad_(True)
    lb = torch.randn(19, 125, generator=g).mul(3).requires_grad_(True)
    clss = torch.randint(0, 125, (19,), generator=g)
    ce_loss = torch.nn.CrossEntropyLoss()                               # reference :160
    l_ims, l_poss = ce_loss(la, clss), ce_loss(lb, clss)                # :240-241
    (1.7 * (l_ims + l_poss)).backward()                                 # :242 (+ an upstream scale)
    out.update(ce_a=la.detach().numpy(), ce_b=lb.detach().numpy(), ce_t=clss.numpy(),
               ce_losses=np.array([(l_ims + l_poss).item(), l_ims.item(), l_poss.item()], dtype=np.float32),
               ce_da=la.grad.numpy(), ce_db=lb.grad.numpy())
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    OUT.mkdir(parents=True, exist_ok=True)
    ContrastiveLoss = load_reference_contrastive()
    np.savez_compressed(OUT / "losses.npz", **golden_losses(ContrastiveLoss))
    np.savez_compressed(OUT / "retrieval.npz", **golden_retrieval())
    np.savez_compressed(OUT / "producer_consumer.npz", **golden_producer_consumer())
    np.savez_compressed(OUT / "autocast.npz", **golden_autocast(ContrastiveLoss))
    for f in sorted(OUT.glob("*.npz")):
        print(f, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
