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
