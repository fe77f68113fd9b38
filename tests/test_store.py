"""Gallery store: file format, streamed scan, DDP-wide gather.
CPU tests cover the format and the host logic; GPU tests check that a streamed / reloaded gallery
returns bit for bit what the resident search returns."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import imageretrievalresearch_b200 as irr
from imageretrievalresearch_b200 import store as st
from oracle import reference_path as ref
from oracle import synthetic


# ------------------------------------------------------------------------------------ CPU: format
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_file_roundtrip_and_layout(tmp_path, dtype):
    N, D = 1000, 72 if dtype == torch.float32 else 64
    emb = torch.randn(N, D).to(dtype)
    lab = torch.arange(N) % 7
    p = tmp_path / "g.irrg"
    with irr.GalleryWriter(p, D, dtype, eps=1e-6, with_labels=True) as w:
        for lo, hi in irr.block_ranges(N, 333):                   # ragged appends
            w.append(emb[lo:hi], lab[lo:hi])
    s = irr.GalleryStore(p)
    assert (s.rows, s.dim, s.dtype, s.has_labels, s.has_inv_norm) == (N, D, dtype, True, False)
    assert torch.equal(s.embeddings(), emb) and torch.equal(s.labels(), lab)
    assert torch.equal(s.embeddings(10, 20), emb[10:20])
    # the embedding block is plain row-major at offset 4096: what a TMA descriptor / mmap reads
    raw = np.fromfile(p, dtype=np.uint8, offset=st.HEADER_BYTES, count=N * D * emb.element_size())
    assert raw.tobytes() == st._as_numpy_rows(emb).tobytes()
    assert os.path.getsize(p) % 8 == 0
    with pytest.raises(IndexError):
        s.embeddings(5, N + 1)


def test_rejects_foreign_truncated_and_misaligned(tmp_path):
    p = tmp_path / "bad.irrg"
    p.write_bytes(b"\0" * 8192)
    with pytest.raises(ValueError, match="magic"):
        irr.GalleryStore(p)
    q = tmp_path / "ok.irrg"
    irr.write_gallery(q, torch.randn(50, 64))
    data = q.read_bytes()
    (tmp_path / "cut.irrg").write_bytes(data[:-100])
    with pytest.raises(ValueError, match="truncated"):
        irr.GalleryStore(tmp_path / "cut.irrg")
    with pytest.raises(ValueError, match="16 bytes"):
        irr.GalleryWriter(tmp_path / "x.irrg", 6, torch.float32)
    with pytest.raises(TypeError):
        irr.GalleryWriter(tmp_path / "x.irrg", 64, torch.int8)
    w = irr.GalleryWriter(tmp_path / "y.irrg", 64, torch.float32, with_labels=True)
    with pytest.raises(ValueError, match="labels"):
        w.append(torch.randn(3, 64))
    w.close()
    assert irr.GalleryStore(tmp_path / "y.irrg").rows == 0


def test_writer_that_fails_midway_leaves_no_gallery(tmp_path):
    """An append that raises inside the `with` block must not seal a short file: the reader would
    accept it and later searches would silently miss rows."""
    p = tmp_path / "partial.irrg"
    with pytest.raises(ValueError, match="expected"):
        with irr.GalleryWriter(p, 64, torch.float32) as w:
            w.append(torch.randn(10, 64))
            w.append(torch.randn(10, 32))             # wrong width: raises part-way through
    assert not p.exists()
    # the same bytes sealed by nobody are rejected, too (magic still zero)
    w = irr.GalleryWriter(p, 64, torch.float32)
    w.append(torch.randn(10, 64))
    w._f.flush()
    with pytest.raises(ValueError, match="magic"):
        irr.GalleryStore(p)
    w.abort()
    assert not p.exists()
    w.abort()                                          # idempotent
    with pytest.raises(RuntimeError, match="closed"):
        w.append(torch.randn(1, 64))


def test_block_ranges():
    assert irr.block_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert irr.block_ranges(0, 4) == []
    assert irr.block_ranges(4, 4) == [(0, 4)]
    with pytest.raises(ValueError):
        irr.block_ranges(4, 0)


def test_streamed_gallery_needs_cuda():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.StreamedGallery(torch.randn(8, 64), 4, "cpu")


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.full((3, 8), float(rank)) + torch.arange(3)[:, None]
    lab = torch.arange(3) + 10 * rank
    g, l, first = irr.gather_embeddings(local, labels=lab)
    ok = (first == 3 * rank and g.shape == (3 * world, 8)
          and all(torch.equal(g[3 * r:3 * r + 3], torch.full((3, 8), float(r)) + torch.arange(3)[:, None])
                  for r in range(world))
          and l.tolist() == [j + 10 * r for r in range(world) for j in range(3)])
    out[rank] = ok
    dist.destroy_process_group()


def test_gather_embeddings_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(world, 29631, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


# ------------------------------------------------------------------------------------ GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,k", [(torch.bfloat16, 3), (torch.float32, 3), (torch.bfloat16, 150),
                                     (torch.float16, 3)])
@pytest.mark.parametrize("pinned", [True, False])
def test_streamed_equals_resident(dtype, k, pinned):
    N, D, Q = 50_011, 256, 70
    q, gal = synthetic.tied_gallery(N, D, Q, dtype=dtype)       # exact ties across block borders
    host = gal.pin_memory() if pinned else gal
    want = irr.cosine_topk(q.cuda(), gal.cuda(), k)
    sg = irr.StreamedGallery(host, 7_001, "cuda", buffers=3 if pinned else 2)
    assert len(sg.blocks) == 8
    for _ in range(2):                                          # second scan reuses the buffers
        got = sg.search(q.cuda(), k)
        assert torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)


@pytest.mark.gpu
def test_store_load_stream_and_labels_on_device(tmp_path):
    N, D, Q, k = 30_000, 1536, 33, 3
    qs, gal, planted = synthetic.planted_gallery(N, D, Q, k, seed=3, dtype=torch.bfloat16)
    lab = torch.arange(N) % 50
    p = tmp_path / "g.irrg"
    with irr.GalleryWriter(p, D, torch.bfloat16, with_labels=True) as w:
        for lo, hi in irr.block_ranges(N, 4096):
            w.append(gal[lo:hi].cuda(), lab[lo:hi])            # CUDA batches: norms are stored
    s = irr.GalleryStore(p)
    assert s.has_inv_norm and torch.equal(s.embeddings(), gal)
    want = irr.Gallery(gal.cuda()).search(qs.cuda(), k)
    assert torch.equal(want.indices.cpu(), planted)
    assert torch.equal(s.inv_norm().cuda(), irr.Gallery(gal.cuda()).inv_norm)
    for got in (s.load("cuda").search(qs.cuda(), k), s.stream("cuda", block_rows=7000).search(qs.cuda(), k)):
        assert torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)
    part = s.load("cuda", 10_000, 20_000)                      # a row range keeps global indices
    r = part.search(qs.cuda(), k)
    full = irr.cosine_topk(qs.cuda(), gal[10_000:20_000].cuda(), k)
    assert torch.equal(r.indices, full.indices + 10_000)
    hits = irr.topk_hits(want.indices, lab[planted[:, 0]].cuda(), s.labels().cuda())
    assert hits.tolist() == [Q, Q]


@pytest.mark.gpu
def test_gallery_builder_grows_and_matches_one_shot():
    N, D, Q, k = 5_000, 256, 40, 3
    q, gal = synthetic.tied_gallery(N, D, Q, dtype=torch.bfloat16)
    lab = torch.arange(N) % 13
    b = irr.GalleryBuilder(D, torch.bfloat16, "cuda", capacity=64)          # forces several regrowths
    for lo, hi in irr.block_ranges(N, 777):
        b.append(gal[lo:hi].cuda() if lo % 2 else gal[lo:hi], lab[lo:hi])   # device and host batches
        if hi == 1554:                                                     # searchable while growing
            part = b.gallery().search(q.cuda(), k)
            want = irr.Gallery(gal[:hi].cuda()).search(q.cuda(), k)       # same cached-norm path
            assert torch.equal(part.indices, want.indices) and torch.equal(part.values, want.values)
    full = irr.Gallery(gal.cuda())
    got, want = b.gallery().search(q.cuda(), k), full.search(q.cuda(), k)
    assert b.rows == N and torch.equal(b.labels.cpu(), lab)
    assert torch.equal(got.indices, want.indices) and torch.equal(got.values, want.values)
    assert torch.equal(b.gallery().inv_norm, full.inv_norm)


def test_gallery_builder_needs_cuda():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irr.GalleryBuilder(64, torch.float32, "cpu")
