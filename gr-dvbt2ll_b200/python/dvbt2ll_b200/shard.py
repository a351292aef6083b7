"""Multi-GPU sharding of the modulator chain (one process per GPU).

T2 frames are independent once BB framing is done (the only carried state -- packet phase, CRC-8 of the
packet in flight, in-band phase, L1 FRAME_IDX -- is a closed-form function of the stream position), so
work is partitioned with NO data-path collective: whole channels per rank when there are at least as
many channels as ranks (config 5), otherwise contiguous runs of T2 frames of a channel.  The only
exchange step is the ordered reassembly of the finished frames on one rank, an NCCL gather.
"""
import numpy as np


def channels_for_rank(n_channels, world, rank):
    """Contiguous block of channel indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_channels, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def frames_for_rank(n_frames, world, rank):
    """(first_frame, count) of the contiguous run of T2 frames of ONE channel owned by `rank`."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def ts_slice_for_frames(chain, first_frame, count):
    """Byte range [lo, hi) of a channel's TS that chain.run_host(row, 1, count, first_frame) must be handed for frames
    [first_frame, first_frame + count): the history bytes (187 in normal input mode once the stream has started: the
    CRC-8 that replaces the first sync byte covers the packet in flight) followed by the frames' TS bytes.  The bytes
    per frame come from the chain (chain.ts_bytes): in high-efficiency mode they depend on the stream position.
    `chain` may also be the constant bytes per T2 frame of a normal-mode stream (int)."""
    if isinstance(chain, (int, np.integer)):
        n = int(chain)
        lo = first_frame * n
        return max(0, lo - 187), lo + count * n
    lo = chain.ts_bytes(0, first_frame) if first_frame > 0 else 0
    return lo - chain.history_bytes(first_frame), lo + chain.ts_bytes(first_frame, count)


def slot_layout(parts_per_rank, bytes_per_part):
    """Ordered reassembly layout of one step: rank r's parts follow rank r-1's.  Returns (offsets, sizes, slot_bytes)."""
    offs, sizes, o = [], [], 0
    for n in parts_per_rank:
        offs.append(o)
        sizes.append(n * bytes_per_part)
        o += n * bytes_per_part
    return offs, sizes, o


def gather_frames(local, dst=0, group=None):
    """Ordered reassembly on rank `dst`: returns the concatenation over ranks (rank order = channel / frame
    order by construction of the partitions above) on dst, None elsewhere.  `local` is a torch tensor;
    NCCL on GPUs, gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    if local.is_complex():                      # collectives move real pairs (gloo has no complex types)
        out = gather_frames(torch.view_as_real(local.contiguous()), dst=dst, group=group)
        return None if out is None else torch.view_as_complex(out.reshape(-1, 2))
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    flat = local.reshape(-1)
    if len(set(sizes)) == 1:
        bufs = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
        dist.gather(flat, bufs, dst=dst, group=group)
        return torch.cat(bufs) if rank == dst else None
    # ragged: pad to the maximum, trim on the root
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=flat.dtype, device=flat.device)
    pad[:flat.numel()] = flat
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)])


def collective_warmup(step, sync, any_rank, min_steps, min_seconds, chunk=8):
    """Warm-up for a job whose ranks are coupled only through the reassembly ring's device-side counters: every rank must
    issue the SAME steps between two host synchronisations (a rank that runs ahead of the root by more than the ring's
    depth and then synchronises waits for slot releases the root only issues with its own later steps).  So the steps go
    in chunks of `chunk`, each followed by `sync()`, and the decision to stop is collective: `any_rank(flag)` must return
    the OR of the ranks' flags (an all-reduce).  Returns the number of steps issued -- the same on every rank."""
    import time
    t0, n = time.perf_counter(), 0
    while True:
        for _ in range(chunk):
            step()
        n += chunk
        sync()
        done = n >= min_steps and time.perf_counter() - t0 >= min_seconds
        if any_rank(bool(done)):
            return n
