"""Host-side sharding of the frame-analysis path across the GPUs of one box.

The reference is single-process ("Parallelism strategies": none); this is
the NEW data-parallel layer.  Two decompositions, neither needs a data-path collective:

  * independent clips  -> contiguous clip ranges per rank (clip_range)
  * one long stream    -> hop-aligned chunks, each loaded with a one-window halo of
                          n - hop samples (chunk_plan); frames of chunk c are the global
                          frames [first_frame, first_frame + n_frames)

The only collective is one all-gather of the fixed-size per-clip summaries
(gather_summaries), issued through torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass


def clip_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: first (n_clips % world) ranks get one extra clip."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Chunk:
    start: int        # first sample of the chunk (hop aligned)
    length: int       # samples to load, including the n - hop halo at the end
    first_frame: int  # global index of the chunk's first frame
    n_frames: int


def total_frames(total_len: int, n: int, hop: int) -> int:
    return 0 if total_len < n else (total_len - n) // hop + 1


def chunk_plan(total_len: int, n: int, hop: int, n_chunks: int) -> list[Chunk]:
    """Split the frames of one stream into n_chunks contiguous frame ranges.

    Chunk c covers frames [f0, f1) and therefore samples [f0*hop, (f1-1)*hop + n): the
    next chunk's first n - hop samples are this chunk's last ones (the window halo).
    Stateless outputs (spectra, energy, centroid) of the chunks concatenate to exactly
    the unchunked result; time-recurrent ones restart at each chunk boundary.
    """
    T = total_frames(total_len, n, hop)
    if n_chunks <= 0:
        raise ValueError("n_chunks must be positive")
    out = []
    for c in range(n_chunks):
        f0, cnt = clip_range(T, c, n_chunks)
        if cnt == 0:
            continue
        out.append(Chunk(start=f0 * hop, length=(cnt - 1) * hop + n, first_frame=f0, n_frames=cnt))
    return out


def uniform_chunks(total_len: int, n: int, hop: int, frames_per_chunk: int):
    """Equal-sized chunks expressed as overlapping clips for aa_analyze_*: returns
    (n_chunks, clip_len, clip_stride, tail) where the first n_chunks chunks have
    frames_per_chunk frames each and `tail` frames remain (to be run as one more clip)."""
    T = total_frames(total_len, n, hop)
    n_chunks = T // frames_per_chunk
    clip_len = (frames_per_chunk - 1) * hop + n
    clip_stride = frames_per_chunk * hop
    return n_chunks, clip_len, clip_stride, T - n_chunks * frames_per_chunk


def gather_summaries(local, group=None):
    """All-gather per-clip summary records (torch uint8/structured bytes tensor [n_local, 32]).

    Every rank must hold the same number of clips (pad with zero records otherwise).
    Returns a tensor [world * n_local, 32] in rank order.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
