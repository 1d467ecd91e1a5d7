"""CPU tests of the multi-GPU host logic: clip ranges, chunk-with-halo plans, and the
summary all-gather over gloo with world_size 2."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

import signals

sh = importlib.import_module("audio-analyzer-rs_b200.sharding")


def test_clip_range_is_a_partition():
    for n in (0, 1, 7, 1024, 65536):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                s, c = sh.clip_range(n, r, world)
                got.extend(range(s, s + c))
            assert got == list(range(n))
            counts = [sh.clip_range(n, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def test_chunk_plan_covers_every_frame_once():
    n, hop = 2048, 512
    for total in (2048, 100000, 172800):
        T = sh.total_frames(total, n, hop)
        for k in (1, 2, 4, 8):
            plan = sh.chunk_plan(total, n, hop, k)
            frames = []
            for ch in plan:
                assert ch.start == ch.first_frame * hop and ch.start % hop == 0
                assert ch.length == (ch.n_frames - 1) * hop + n
                assert ch.start + ch.length <= total
                frames.extend(range(ch.first_frame, ch.first_frame + ch.n_frames))
            assert frames == list(range(T))
            for a, b in zip(plan, plan[1:]):
                # one-window halo: the next chunk starts n - hop samples before this one ends
                assert a.start + a.length - b.start == n - hop


def test_chunked_oracle_stateless_outputs_are_exact(O):
    """cfg4 contract on the CPU side: spectra / energy / centroid of hop-aligned chunks with a
    window halo concatenate to exactly the unchunked result."""
    x = signals.chord_vibrato(0xA0D14, 48000.0, 60000)
    n, hop = 2048, 512
    cfg = O.make_config(n, hop, 48000.0)
    full = O.analyze_clip(cfg, x)
    mags, energy, cent = [], [], []
    for ch in sh.chunk_plan(len(x), n, hop, 4):
        r = O.analyze_clip(cfg, x[ch.start: ch.start + ch.length])
        assert r["T"] == ch.n_frames
        mags.append(r["mags"])
        energy.append(r["features"]["energy"])
        cent.append(r["features"]["centroid"])
    assert np.array_equal(np.concatenate(mags), full["mags"])
    assert np.array_equal(np.concatenate(energy), full["features"]["energy"])
    assert np.array_equal(np.concatenate(cent), full["features"]["centroid"])


def test_uniform_chunks():
    nch, clen, stride, tail = sh.uniform_chunks(172800, 2048, 512, 64)
    T = sh.total_frames(172800, 2048, 512)
    assert nch * 64 + tail == T and clen == 63 * 512 + 2048 and stride == 64 * 512


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, cnt = sh.clip_range(n_clips, rank, world)
    local = torch.zeros((cnt, 32), dtype=torch.uint8)
    for i in range(cnt):
        local[i, :] = (start + i) % 251
    allv = sh.gather_summaries(local)
    q.put((rank, allv[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_summary_all_gather_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_clips, world = 10, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [i % 251 for i in range(n_clips)]
    assert res[0] == want and res[1] == want


def test_chunk_plan_with_warmup():
    n, hop, total = 2048, 512, 172800
    T = sh.total_frames(total, n, hop)
    for k in (1, 3, 8):
        for w in (0, 5, 64, 10 ** 6):
            plan = sh.chunk_plan(total, n, hop, k, warmup_frames=w)
            frames = []
            for ch in plan:
                assert ch.warmup == min(w, ch.first_frame)              # the first chunk has no past to warm up on
                assert ch.start == (ch.first_frame - ch.warmup) * hop
                assert ch.length == (ch.n_frames + ch.warmup - 1) * hop + n
                assert ch.start >= 0 and ch.start + ch.length <= total
                frames.extend(range(ch.first_frame, ch.first_frame + ch.n_frames))
            assert frames == list(range(T))
    with pytest.raises(ValueError):
        sh.chunk_plan(total, n, hop, 4, warmup_frames=-1)


def test_exact_schedule_is_a_topological_partition():
    for n_streams, n_chunks, world in ((1, 7, 2), (3, 5, 2), (8, 16, 8), (2, 3, 4)):
        seen = {}
        for r in range(world):
            items = sh.exact_schedule(n_streams, n_chunks, r, world)
            waves = [s + c for s, c in items]
            assert waves == sorted(waves)
            for pos, (s, c) in enumerate(items):
                assert sh.owner(c, world) == r and (s, c) not in seen
                seen[(s, c)] = (r, s + c)
        assert len(seen) == n_streams * n_chunks
        # the predecessor of every item sits on an earlier wavefront
        for (s, c), (_, wave) in seen.items():
            if c:
                assert seen[(s, c - 1)][1] == wave - 1


def _fake_run_chunk(x, frames_per_chunk, out):
    """A stand-in analyzer with the structure of the real one: per-'bin' f32 recurrences with a data-dependent
    branch, one record per frame, state = the recurrent values."""
    import torch

    def run(s, c, state):
        for f in range(c * frames_per_chunk, (c + 1) * frames_per_chunk):
            v = x[s, f]
            upd = torch.where(v > state, state + 0.1 * (v - state), state + 0.02 * (v - state))
            state.copy_(upd)
            out[s, f] = state.sum()
    return run


def _chain_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_streams, n_chunks, fpc, bins = 3, 5, 4, 16
    g = torch.Generator().manual_seed(7)
    x = torch.rand((n_streams, n_chunks * fpc, bins), generator=g)
    out = torch.zeros((n_streams, n_chunks * fpc))
    send, recv = sh.torch_send_recv()
    log = sh.chain_chunks(n_streams, n_chunks, rank, world, _fake_run_chunk(x, fpc, out),
                          lambda: torch.zeros(bins), send, recv)
    dist.all_reduce(out)            # every frame was produced by exactly one rank
    q.put((rank, out.tolist(), log))
    dist.barrier()
    dist.destroy_process_group()


def test_exact_chain_over_gloo_world2_equals_the_unchunked_run():
    """cfg4 exact mode, host logic: chunks of several streams alternate between two ranks, the analyzer state is
    handed over as a message, and the records equal the unchunked single-process run bit for bit."""
    import torch
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (o, lg) for r, o, lg in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_streams, n_chunks, fpc, bins = 3, 5, 4, 16
    g = torch.Generator().manual_seed(7)
    x = torch.rand((n_streams, n_chunks * fpc, bins), generator=g)
    want = torch.zeros((n_streams, n_chunks * fpc))
    run = _fake_run_chunk(x, n_chunks * fpc, want)
    for s in range(n_streams):
        run(s, 0, torch.zeros(bins))               # one chunk = the whole stream
    assert res[0][0] == want.tolist() and res[1][0] == want.tolist()
    # every chunk after the first got its state from the other rank
    for r in (0, 1):
        for s, c, src in res[r][1]:
            assert src == (None if c == 0 else 1 - r)
    # single rank: the same driver keeps every state local
    out1 = torch.zeros_like(want)
    log = sh.chain_chunks(n_streams, n_chunks, 0, 1, _fake_run_chunk(x, fpc, out1), lambda: torch.zeros(bins))
    assert out1.tolist() == want.tolist() and all(src is None for _, _, src in log)
