#!/usr/bin/env python
"""BASELINE configs[3] and [4] under torchrun (one process per GPU):

  cfg4  1-hour synthetic stream @ 48 kHz, 2048-pt / hop 512, split into hop-aligned chunks with a
        one-window halo; rank r analyses chunks [r*C/W, (r+1)*C/W).  Rank 0 also runs the first
        minutes unchunked and checks that the stateless outputs (energy, centroid) are bit-identical.
  cfg5  65536 synthetic 10 s clips @ 44.1 kHz, 2048-pt / hop 512, all features, sharded by clip range;
        per-clip summaries all-gathered over NCCL and checked on every rank.

Rank 0 prints one JSON line per case (device time = max over ranks, CUDA events)."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
aa = importlib.import_module("audio-analyzer-rs_b200")
sh = importlib.import_module("audio-analyzer-rs_b200.sharding")


def max_over_ranks(ms, dev):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cfg4(rank, world, dev, seconds=3600.0):
    n, hop, sr = 2048, 512, 48000.0
    total = int(seconds * sr)
    frames_per_chunk = 256
    stream = torch.empty(1, total, device=dev)
    aa.synth_clips_device(stream.data_ptr(), 1, total, total, sr, 0xA0D14)      # every rank: same stream
    nch, clen, stride, tail = sh.uniform_chunks(total, n, hop, frames_per_chunk)
    c0, cnt = sh.clip_range(nch, rank, world)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    feat = torch.zeros(cnt * frames_per_chunk, 96, device=dev, dtype=torch.uint8)
    summ = torch.zeros(cnt, 32, device=dev, dtype=torch.uint8)
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        an.analyze_device(stream.data_ptr() + 4 * c0 * stride, cnt, clen, stride, features=feat.data_ptr(),
                          summaries=summ.data_ptr(), stream=s.cuda_stream)
    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(s)
    run()
    e1.record(s)
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    # equal chunk counts per rank are needed by the all-gather: pad to the maximum
    cmax = sh.clip_range(nch, 0, world)[1]
    pad = torch.zeros(cmax, 32, device=dev, dtype=torch.uint8)
    pad[:cnt] = summ
    allsum = sh.gather_summaries(pad)
    ok = None
    if rank == 0:
        # unchunked reference of the first 60 s on one CTA: stateless outputs must be identical
        ref_len = int(60 * sr)
        T = an.num_frames(ref_len)
        rf = torch.zeros(T, 96, device=dev, dtype=torch.uint8)
        an.analyze_device(stream.data_ptr(), 1, ref_len, ref_len, features=rf.data_ptr(), stream=s.cuda_stream)
        torch.cuda.synchronize()
        a = rf.cpu().numpy().view(aa.FEATURES_DTYPE).reshape(-1)
        nchk = min(cnt, T // frames_per_chunk)
        b = feat[: nchk * frames_per_chunk].cpu().numpy().view(aa.FEATURES_DTYPE).reshape(-1)
        m = nchk * frames_per_chunk
        ok = bool(np.array_equal(a["energy"][:m], b["energy"]) and np.array_equal(a["centroid"][:m], b["centroid"]))
        agree = float((a["n_pitches"][:m] == b["n_pitches"]).mean())
        frames = nch * frames_per_chunk
        print(json.dumps({"case": "cfg4 1-hour stream, hop-aligned chunks with window halo", "n_gpus": world,
                          "chunks": nch, "frames_per_chunk": frames_per_chunk, "frames": frames, "ms": ms,
                          "frames_per_s": frames / ms * 1e3, "audio_s_per_s": frames * hop / sr / ms * 1e3,
                          "stateless_outputs_bit_identical_to_unchunked": ok,
                          "pitch_count_agreement_with_unchunked": agree,
                          "gathered_summaries": int(allsum.shape[0])}), flush=True)
    return ok


def cfg5(rank, world, dev, total_clips=65536):
    n, hop, sr = 2048, 512, 44100.0
    clip_len = 441000
    c0, cnt = sh.clip_range(total_clips, rank, world)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    T = an.num_frames(clip_len)
    clips = torch.empty(cnt, clip_len, device=dev)
    aa.synth_clips_device(clips.data_ptr(), cnt, clip_len, clip_len, sr, 0xA0D15 + c0)
    feat = torch.empty(cnt * T, 96, device=dev, dtype=torch.uint8)
    stab = torch.empty(cnt * T, 136, device=dev, dtype=torch.uint8)
    summ = torch.zeros(cnt, 32, device=dev, dtype=torch.uint8)
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        an.analyze_device(clips.data_ptr(), cnt, clip_len, clip_len, features=feat.data_ptr(),
                          stable=stab.data_ptr(), summaries=summ.data_ptr(), stream=s.cuda_stream)
        return sh.gather_summaries(summ)
    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(s)
    allsum = run()
    e1.record(s)
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    h = allsum.cpu().numpy().view(aa.SUMMARY_DTYPE).reshape(-1)
    good = bool(h.shape[0] == total_clips and (h["n_frames"] == T).all() and (h["n_pitched"] > 0).all())
    # my own rows of the gathered table are my local summaries
    mine = summ.cpu().numpy().view(aa.SUMMARY_DTYPE).reshape(-1)
    good = good and h[c0:c0 + cnt].tobytes() == mine.tobytes()
    flag = torch.tensor([1 if good else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        frames = total_clips * T
        print(json.dumps({"case": "cfg5 65536 clips x 10 s @ 44.1 kHz, 2048/512, all features, features-only output",
                          "n_gpus": world, "clips_per_gpu": cnt, "frames": frames, "ms": ms,
                          "frames_per_s": frames / ms * 1e3, "audio_s_per_s": total_clips * 10.0 / ms * 1e3,
                          "summary_table_consistent_on_all_ranks": bool(flag.item())}), flush=True)


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    aa.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    clips5 = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    if which in ("cfg4", "both"):
        cfg4(rank, world, dev)
    if which in ("cfg5", "both"):
        cfg5(rank, world, dev, clips5)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
