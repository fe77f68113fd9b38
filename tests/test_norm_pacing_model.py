"""Model of the in-kernel gallery-norm producers of the CTA-pair top-k kernel
(csrc/cosine_topk_bf16.cu, NORMS_INSIDE): the slot enumeration and the pacing rule restated in
Python, checked on the host for shapes the GPU tests do not reach.

  * every octet of every gallery tile is produced exactly once, by exactly one producer warp,
    and nothing outside the gallery is touched;
  * a producer's slots come in non-decreasing order of their timeline key;
  * with the pacing rule (a producer never runs more than `ahead` tiles ahead of its own CTA's
    epilogue) and consumers that wait for their tile's counter, every interleaving terminates.
"""
import random

import pytest

BLOCK_M, BLOCK_N, OCTETS = 128, 256, 32
SMS = 148


def plan_pair(Q, N, D=1536, k=3):
    """make_plan_pair(): chunks short enough that a wave's tiles stay in L2 from 12 pairs on (k <= 4)"""
    clusters = SMS // 2
    m_pairs = (Q + 2 * BLOCK_M - 1) // (2 * BLOCK_M)
    n_tiles = max((N + BLOCK_N - 1) // BLOCK_N, 1)
    best, best_cost = 1, 1e300
    max_tpc = min(n_tiles, 64)
    if m_pairs >= 12 and k <= 4 and n_tiles * m_pairs * D >= 512 * clusters * 1536:
        chunks_per_wave = (clusters + m_pairs - 1) // m_pairs
        max_tpc = min(max_tpc, max((32 << 20) // (chunks_per_wave * BLOCK_N * D * 2), 4))
    for tpc in range(1, max_tpc + 1):
        chunks = (n_tiles + tpc - 1) // tpc
        waves = (chunks * m_pairs + clusters - 1) // clusters
        cost = waves * (tpc + 0.35)
        if cost < best_cost - 1e-9:
            best, best_cost = tpc, cost
    n_chunks = (n_tiles + best - 1) // best
    num_clusters = min(n_chunks * m_pairs, clusters)
    return m_pairs, n_tiles, best, n_chunks, num_clusters


def slots_of(nw, NW, m_pairs, n_tiles, tpc, n_chunks, C, N):
    """The producer loop of warp `nw`: yields (key, tile, octet, rows) in program order."""
    spc = tpc * OCTETS
    o = nw
    while o < spc * n_chunks:
        cq = o // spc
        w = cq * m_pairs // C
        c_lo = (w * C + m_pairs - 1) // m_pairs
        c_hi = min(((w + 1) * C + m_pairs - 1) // m_pairs, n_chunks)
        nch = c_hi - c_lo
        rem = o - c_lo * spc
        j = rem // (nch * OCTETS)
        r2 = rem - j * (nch * OCTETS)
        chunk, octet = c_lo + r2 // OCTETS, r2 % OCTETS
        assert c_lo <= cq < c_hi and 0 <= j < tpc and c_lo <= chunk < c_hi
        tile = chunk * tpc + j
        o += NW
        if tile >= n_tiles:
            continue
        row0 = tile * BLOCK_N + octet * 8
        if row0 >= N:
            continue
        yield w * tpc + j, tile, octet, min(8, N - row0)


SHAPES = [(4096, 1_000_000), (8192, 1_250_000), (4096, 125_000), (600, 3000), (640, 3000),
          (768, 1_000_000), (1024, 50_000), (20_000, 100_000), (40_000, 7_777), (700, 255),
          (513, 257), (2048, 999_999)]


@pytest.mark.parametrize("Q,N", SHAPES)
def test_every_octet_exactly_once_in_key_order(Q, N):
    m_pairs, n_tiles, tpc, n_chunks, C = plan_pair(Q, N)
    NW = 2 * C * 4
    seen = {}
    for nw in range(NW):
        last = -1
        for key, tile, octet, rows in slots_of(nw, NW, m_pairs, n_tiles, tpc, n_chunks, C, N):
            assert key >= last
            last = key
            assert (tile, octet) not in seen
            seen[(tile, octet)] = rows
    rows_per_tile = {}
    for (tile, _), rows in seen.items():
        rows_per_tile[tile] = rows_per_tile.get(tile, 0) + rows
    assert sorted(rows_per_tile) == list(range(n_tiles))
    assert all(rows_per_tile[t] == min(BLOCK_N, N - t * BLOCK_N) for t in range(n_tiles))


@pytest.mark.parametrize("Q,N", [(600, 3000), (1024, 50_000), (40_000, 7_777), (4096, 40_000), (513, 257)])
@pytest.mark.parametrize("ahead", [0, 1, 2])
def test_paced_producers_and_waiting_consumers_always_finish(Q, N, ahead):
    m_pairs, n_tiles, tpc, n_chunks, C = plan_pair(Q, N)
    total_units = m_pairs * n_chunks
    NW = 2 * C * 4
    need = [min(BLOCK_N, N - t * BLOCK_N) for t in range(n_tiles)]
    for seed in range(3):
        rng = random.Random(seed)
        done = [0] * n_tiles
        # consumers: one epilogue per CTA (2 per cluster), walking (unit, tile) pairs
        cons = []
        for cta in range(2 * C):
            cl = cta // 2
            walk = [((u // C) * tpc + (t - (u // m_pairs) * tpc), t)
                    for u in range(cl, total_units, C)
                    for t in range((u // m_pairs) * tpc, min((u // m_pairs) * tpc + tpc, n_tiles))]
            cons.append(walk)
        cpos = [0] * (2 * C)                 # index into the walk
        pos = [0] * (2 * C)                  # published stream position
        prods = [list(slots_of(nw, NW, m_pairs, n_tiles, tpc, n_chunks, C, N)) for nw in range(NW)]
        ppos = [0] * NW
        while True:
            moves = []
            for cta in range(2 * C):
                if cpos[cta] < len(cons[cta]):
                    p, t = cons[cta][cpos[cta]]
                    pos[cta] = p                                   # published before waiting
                    if done[t] >= need[t]:
                        moves.append(("c", cta))
                else:
                    pos[cta] = 0x3fffffff
            for nw in range(NW):
                if ppos[nw] < len(prods[nw]) and prods[nw][ppos[nw]][0] <= pos[nw // 4] + ahead:
                    moves.append(("p", nw))
            if not moves:
                break
            kind, i = rng.choice(moves)
            if kind == "c":
                cpos[i] += 1
            else:
                _, t, _, rows = prods[i][ppos[i]]
                done[t] += rows
                ppos[i] += 1
        assert all(cpos[c] == len(cons[c]) for c in range(2 * C)), "consumers stuck"
        assert all(ppos[n] == len(prods[n]) for n in range(NW)), "producers stuck"
        assert done == need


@pytest.mark.parametrize("Q,N", SHAPES)
def test_workspace_covers_the_pair_plan_and_the_tile_counters(Q, N):
    """irr_cosine_topk_workspace_bytes (callable without a device) is at least what the pair
    kernel's plan needs: cached norms + two partial-list arrays + row floors + per-tile counters."""
    from imageretrievalresearch_b200 import _lib
    lib = _lib.load()
    for k in (1, 3, 10, 16):
        m_pairs, n_tiles, tpc, n_chunks, _ = plan_pair(Q, N, 1536, k)
        parts = n_chunks * Q * k
        need = N * 4 + 2 * parts * 4 + Q * 4 + n_tiles * 4
        got = lib.irr_cosine_topk_workspace_bytes(Q, N, 1536, k, _lib.IRR_BF16)
        assert got >= need, (Q, N, k, got, need)
        assert got <= need + 8 * 256 + max(need // 4, 1 << 16), "workspace far larger than the plan"
