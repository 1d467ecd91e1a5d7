"""Host-side sharding of the frame-analysis path across the GPUs of one box.

The reference is single-process ("Parallelism strategies": none); this is
the NEW data-parallel layer.  Two decompositions, neither needs a data-path collective:

  * independent clips  -> contiguous clip ranges per rank (clip_range)
  * one long stream    -> hop-aligned chunks, each loaded with a one-window halo of
                          n - hop samples (chunk_plan); frames of chunk c are the global
                          frames [first_frame, first_frame + n_frames).  The reference's
                          analyzers carry per-bin floors / trackers forever (stft.rs:209-212,
                          338-363), so the time-recurrent outputs of a chunk depend on the past:
                            - warm-up mode: every chunk starts `warmup_frames` early with a fresh
                              analyzer and discards those frames (embarrassingly parallel, NOT exact:
                              the mismatch against the unchunked run is a measured function of the
                              warm-up length);
                            - exact mode: chunk c+1 starts from the analyzer state block chunk c
                              left (aa_analyze_device_carry), handed from rank to rank as one
                              ~33 KB message (chain_chunks): bit-identical to the unchunked run,
                              serial along one stream, pipelined across streams (exact_schedule).

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
    start: int        # first sample to load (hop aligned; includes the warm-up frames)
    length: int       # samples to load, including the n - hop halo at the end
    first_frame: int  # global index of the chunk's first KEPT frame
    n_frames: int     # kept frames
    warmup: int = 0   # frames analysed before first_frame only to warm the recurrent state up (discarded)


def total_frames(total_len: int, n: int, hop: int) -> int:
    return 0 if total_len < n else (total_len - n) // hop + 1


def chunk_plan(total_len: int, n: int, hop: int, n_chunks: int, warmup_frames: int = 0) -> list[Chunk]:
    """Split the frames of one stream into n_chunks contiguous frame ranges.

    Chunk c covers frames [f0, f1) and therefore samples [f0*hop, (f1-1)*hop + n): the
    next chunk's first n - hop samples are this chunk's last ones (the window halo).
    Stateless outputs (spectra, energy, centroid) of the chunks concatenate to exactly
    the unchunked result; time-recurrent ones restart at each chunk boundary unless the
    chunk is given `warmup_frames` of run-in (it then starts min(warmup_frames, f0) frames
    early and the caller drops the records of those frames) or the state is carried
    (chain_chunks).
    """
    T = total_frames(total_len, n, hop)
    if n_chunks <= 0:
        raise ValueError("n_chunks must be positive")
    if warmup_frames < 0:
        raise ValueError("warmup_frames must not be negative")
    out = []
    for c in range(n_chunks):
        f0, cnt = clip_range(T, c, n_chunks)
        if cnt == 0:
            continue
        w = min(warmup_frames, f0)
        out.append(Chunk(start=(f0 - w) * hop, length=(cnt + w - 1) * hop + n, first_frame=f0, n_frames=cnt,
                         warmup=w))
    return out


def owner(chunk_index: int, world: int) -> int:
    """Round-robin owner of a chunk: consecutive chunks sit on consecutive ranks, so the state of a stream walks
    around the ring of GPUs."""
    return chunk_index % world


def exact_schedule(n_streams: int, n_chunks: int, rank: int, world: int) -> list[tuple[int, int]]:
    """Work items (stream, chunk) of `rank` in exact mode, in the order it must run them.  Item (s, c) needs the
    state of (s, c - 1); items are ordered by the wavefront s + c (then s), which is a topological order of that
    dependency on every rank, so blocking receives cannot deadlock as long as sends do not block (isend).  With
    n_streams >= world every rank has work on every wavefront after the first world - 1."""
    items = [(s, c) for s in range(n_streams) for c in range(n_chunks) if owner(c, world) == rank]
    return sorted(items, key=lambda sc: (sc[0] + sc[1], sc[0]))


def chain_chunks(n_streams: int, n_chunks: int, rank: int, world: int, run_chunk, new_state, send=None, recv=None):
    """Exact mode driver.  run_chunk(stream, chunk, state) analyses one chunk starting from `state` and leaves the
    final analyzer state in it (in place); new_state() returns a fresh (all-zero) state block; send(state, dst,
    tag) must not block (torch.distributed.isend), recv(state, src, tag) blocks.  Returns the hand-off log
    [(stream, chunk, src_rank or None)] for tests.  State blocks handed to a chunk owned by this same rank never
    leave the device."""
    local: dict[tuple[int, int], object] = {}
    pending = []
    log = []
    for s, c in exact_schedule(n_streams, n_chunks, rank, world):
        src = None
        if c == 0:
            state = new_state()
        elif (s, c) in local:
            state = local.pop((s, c))
        else:
            src = owner(c - 1, world)
            state = new_state()
            recv(state, src, s * n_chunks + c)
        run_chunk(s, c, state)
        log.append((s, c, src))
        if c + 1 < n_chunks:
            dst = owner(c + 1, world)
            if dst == rank:
                local[(s, c + 1)] = state
            else:
                pending.append(send(state, dst, s * n_chunks + c + 1))
    for req in pending:
        if req is not None and hasattr(req, "wait"):
            req.wait()
    return log


def torch_send_recv(group=None):
    """(send, recv) over torch.distributed point-to-point (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist

    def send(state, dst, tag):
        return dist.isend(state, dst, group=group, tag=tag)

    def recv(state, src, tag):
        dist.recv(state, src, group=group, tag=tag)

    return send, recv


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
