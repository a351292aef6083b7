"""N>1 host logic on CPU: partitioning of channels / T2 frames over ranks and the ordered gather, with the
gloo backend and world_size 2 (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest

from dvbt2ll_b200 import shard


def test_partitions_cover_everything_in_order():
    for n, w in ((64, 1), (64, 2), (64, 8), (7, 4), (3, 8)):
        got = [c for r in range(w) for c in shard.channels_for_rank(n, w, r)]
        assert got == list(range(n))
        runs = [shard.frames_for_rank(n, w, r) for r in range(w)]
        pos = 0
        for first, count in runs:
            assert first == pos
            pos += count
        assert pos == n
    lo, hi = shard.ts_slice_for_frames(12352, 2, 3)
    assert (lo, hi) == (2 * 12352 - 187, 5 * 12352)
    assert shard.ts_slice_for_frames(12352, 0, 1) == (0, 12352)


def _worker(rank, world, port, n_channels, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.channels_for_rank(n_channels, world, rank)
    # stand-in for the per-channel baseband: channel c -> 5 samples of value c + j*k
    local = torch.tensor([[complex(c, k) for k in range(5)] for c in mine], dtype=torch.complex64).reshape(len(mine), 5)
    out = shard.gather_frames(local, dst=0)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_channels", [4, 5])
def test_ordered_gather_gloo_world2(n_channels):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_channels, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[complex(c, k) for k in range(5)] for c in range(n_channels)], dtype=np.complex64).reshape(-1)
    assert np.array_equal(out, want)
