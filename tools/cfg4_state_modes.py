#!/usr/bin/env python
"""BASELINE configs[3] (cfg4), the stateful outputs: a long synthetic stream @ 48 kHz, 2048-pt / hop 512, in
hop-aligned chunks across the GPUs of one box (one process per GPU under torchrun; also runs on one GPU).

  unchunked   rank 0 analyses the whole stream as ONE clip (the reference's own reading: an analyzer that never
              forgets, stft.rs:209-212) -- the truth both modes are compared with.
  exact       chunk c+1 starts from the analyzer state block chunk c left (aa_analyze_device_carry); consecutive
              chunks sit on consecutive ranks and the block travels as one NCCL point-to-point message
              (sharding.chain_chunks).  Must be byte-identical to the unchunked run.  Serial along one stream; with
              S >= world independent streams the ranks pipeline (exact_schedule) -- both are timed.
  warm-up     every chunk starts W frames early with a fresh analyzer and drops those frames (embarrassingly
              parallel, not exact): mismatching frames against the unchunked run and the boundary mismatch window
              (how far behind a chunk start the last mismatch sits) as a function of W -- the table of DESIGN.md 4.

Rank 0 prints one JSON line.  usage: cfg4_state_modes.py [seconds=3600] [frames_per_chunk=4096]"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
aa = importlib.import_module("audio-analyzer-rs_b200")
sh = importlib.import_module("audio-analyzer-rs_b200.sharding")


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    fpc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    aa.set_device(local)
    cs = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(cs)
    s = cs.cuda_stream

    n, hop, sr = 2048, 512, 48000.0
    total = int(seconds * sr)
    total -= total % 4
    cfg = aa.Config(n=n, sample_rate=sr)
    an = aa.Analyzer(cfg)
    T = an.num_frames(total)
    n_chunks = (T + fpc - 1) // fpc
    plan = [sh.Chunk(start=c * fpc * hop, length=(min(fpc, T - c * fpc) - 1) * hop + n, first_frame=c * fpc,
                     n_frames=min(fpc, T - c * fpc)) for c in range(n_chunks)]
    n_streams = max(world, 1)
    # stream 0 is the cfg4 stream (seed 0xA0D14); the others (pipelining demo) are further seeds
    xs = torch.empty(n_streams, total, device=dev)
    aa.synth_clips_device(xs.data_ptr(), n_streams, total, total, sr, 0xA0D14)
    torch.cuda.synchronize()
    nfl = an.state_floats

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather_sum(t):          # every frame is written by exactly one rank, the others hold zeros
        if world > 1:
            v = t.view(torch.int32)
            dist.all_reduce(v)
        return t

    # ---- unchunked truth (every rank computes it: it is one CTA's work and keeps the comparison local) ------
    ref_feat = torch.zeros(T, 96, device=dev, dtype=torch.uint8)
    ref_stab = torch.zeros(T, 136, device=dev, dtype=torch.uint8)
    t0 = time.perf_counter()
    an.analyze_device(xs.data_ptr(), 1, total, total, features=ref_feat.data_ptr(), stable=ref_stab.data_ptr(), stream=s)
    torch.cuda.synchronize()
    unchunked_s = time.perf_counter() - t0
    rf = ref_feat.cpu().numpy().view(aa.FEATURES_DTYPE).reshape(-1)
    rs = ref_stab.cpu().numpy().view(aa.STABLE_DTYPE).reshape(-1)

    # ---- exact mode ---------------------------------------------------------------------------------------
    def exact(ns):
        feat = torch.zeros(ns, T, 96, device=dev, dtype=torch.uint8)
        stab = torch.zeros(ns, T, 136, device=dev, dtype=torch.uint8)

        def run_chunk(si, c, state):
            ch = plan[c]
            an.analyze_device_carry(xs[si].data_ptr() + 4 * ch.start, 1, ch.length, (ch.length + 3) & ~3,
                                    state.data_ptr(), features=feat[si].data_ptr() + 96 * ch.first_frame,
                                    stable=stab[si].data_ptr() + 136 * ch.first_frame, stream=s)

        send, recv = sh.torch_send_recv() if world > 1 else (None, None)
        sync_all()
        t0 = time.perf_counter()
        sh.chain_chunks(ns, n_chunks, rank, world, run_chunk, lambda: torch.zeros(nfl, device=dev), send, recv)
        sync_all()
        dt = time.perf_counter() - t0
        return gather_sum(feat), gather_sum(stab), dt

    feat1, stab1, exact1_s = exact(1)
    exact_ok = bool(torch.equal(feat1[0], ref_feat) and torch.equal(stab1[0], ref_stab))
    featS, stabS, exactS_s = exact(n_streams)
    exact_ok = exact_ok and bool(torch.equal(featS[0], ref_feat) and torch.equal(stabS[0], ref_stab))
    del feat1, stab1, featS, stabS

    # ---- warm-up mode: mismatch vs W --------------------------------------------------------------------------
    def cmp_pitch(a, b):
        same = (a["n_pitches"] == b["n_pitches"])
        close = np.isclose(a["pitch"]["freq"], b["pitch"]["freq"], rtol=2e-5, atol=0).all(axis=-1) & \
            np.isclose(a["pitch"]["score"], b["pitch"]["score"], rtol=2e-5, atol=0).all(axis=-1)
        return ~(same & close)

    table = []
    for W in (0, 16, 64, 256, 1024, 4096, 16384):
        feat = torch.zeros(T, 96, device=dev, dtype=torch.uint8)
        stab = torch.zeros(T, 136, device=dev, dtype=torch.uint8)
        c0, cnt = sh.clip_range(n_chunks, rank, world)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cs)
        mine = list(range(c0, c0 + cnt))
        regular = [c for c in mine if plan[c].first_frame >= W and plan[c].n_frames == fpc]
        # chunks with the full warm-up and the full length are overlapping clips of ONE launch (contiguous run)
        if regular:
            ra, rb = regular[0], regular[-1] + 1
            assert regular == list(range(ra, rb))
            k = rb - ra
            length = (fpc + W - 1) * hop + n
            tf = torch.empty(k, fpc + W, 96, device=dev, dtype=torch.uint8)
            ts = torch.empty(k, fpc + W, 136, device=dev, dtype=torch.uint8)
            an.analyze_device(xs.data_ptr() + 4 * (plan[ra].first_frame - W) * hop, k, length, fpc * hop,
                              features=tf.data_ptr(), stable=ts.data_ptr(), stream=s)
            f0 = plan[ra].first_frame
            feat[f0:f0 + k * fpc] = tf[:, W:].reshape(k * fpc, 96)
            stab[f0:f0 + k * fpc] = ts[:, W:].reshape(k * fpc, 136)
        for c in mine:
            if c in regular:
                continue
            ch = plan[c]
            w = min(W, ch.first_frame)
            start = (ch.first_frame - w) * hop
            length = (ch.n_frames + w - 1) * hop + n
            tf = torch.empty(ch.n_frames + w, 96, device=dev, dtype=torch.uint8)
            ts = torch.empty(ch.n_frames + w, 136, device=dev, dtype=torch.uint8)
            an.analyze_device(xs.data_ptr() + 4 * start, 1, length, (length + 3) & ~3, features=tf.data_ptr(),
                              stable=ts.data_ptr(), stream=s)
            feat[ch.first_frame:ch.first_frame + ch.n_frames] = tf[w:]
            stab[ch.first_frame:ch.first_frame + ch.n_frames] = ts[w:]
        e1.record(cs)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gather_sum(feat)
        gather_sum(stab)
        f = feat.cpu().numpy().view(aa.FEATURES_DTYPE).reshape(-1)
        st = stab.cpu().numpy().view(aa.STABLE_DTYPE).reshape(-1)
        bad_p = cmp_pitch(f, rf)
        bad_b = f["burst_count"] != rf["burst_count"]
        bad_fl = f["flags"] != rf["flags"]
        bad_s = (st["n"] != rs["n"]) | ~np.isclose(st["pitch"]["freq"], rs["pitch"]["freq"], rtol=2e-5, atol=0).all(axis=-1)
        off = np.arange(T) % fpc                      # frames since the chunk start
        later = np.arange(T) >= fpc                   # the first chunk has no boundary behind it

        def window(bad):
            b = bad & later
            return int(off[b].max()) + 1 if b.any() else 0

        table.append({"warmup_frames": W, "ms_max_over_ranks": float(ms.item()),
                      "pitch_list_frames": int(bad_p.sum()), "burst_count_frames": int(bad_b.sum()),
                      "onset_flag_frames": int(bad_fl.sum()), "stable_pitch_frames": int(bad_s.sum()),
                      "stateless_equal": bool(np.array_equal(f["energy"], rf["energy"]) and
                                              np.array_equal(f["centroid"], rf["centroid"])),
                      "mismatch_window_frames": {"pitch": window(bad_p), "burst": window(bad_b), "stable": window(bad_s)}})
    if rank == 0:
        print(json.dumps({
            "case": f"cfg4 {seconds:g} s stream @ 48 kHz, 2048-pt / hop 512, {n_chunks} chunks of {fpc} frames",
            "n_gpus": world, "frames": T, "state_block_bytes": 4 * nfl,
            "unchunked_one_cta_s": unchunked_s, "unchunked_frames_per_s": T / unchunked_s,
            "exact": {"byte_identical_to_unchunked": exact_ok, "one_stream_s": exact1_s,
                      "one_stream_frames_per_s": T / exact1_s, "streams_pipelined": n_streams,
                      "pipelined_s": exactS_s, "pipelined_frames_per_s": n_streams * T / exactS_s,
                      "handoffs_per_stream": n_chunks - 1},
            "warmup": table}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
