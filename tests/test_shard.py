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
    assert shard.slot_layout([3, 2, 2], 10) == ([0, 30, 50], [30, 20, 20], 70)


def test_ts_slices_follow_the_chain_in_both_input_modes():
    """The TS range of a run of T2 frames comes from the chain: constant per frame in normal mode (+187 history bytes
    once the stream has started), position dependent and history-free in high-efficiency mode."""
    import dvbt2ll_b200 as T
    from dvbt2ll_b200 import configs as K
    ch = T.Chain(K.resolve("c1"), max_frames=1)
    n = ch.ts_bytes_per_frame
    assert shard.ts_slice_for_frames(ch, 0, 2) == (0, 2 * n)
    assert shard.ts_slice_for_frames(ch, 3, 2) == (3 * n - 187, 5 * n)
    hem = T.Chain(K.resolve(dict(K.CONFIGS["c1"], inputmode=1, inband=1, version=2, fecblocks=7)), max_frames=1)
    sizes = [hem.ts_bytes(f, 1) for f in range(6)]
    assert len(set(sizes)) > 1
    pos = 0
    for f in range(6):
        assert shard.ts_slice_for_frames(hem, f, 1) == (pos, pos + sizes[f])
        pos += sizes[f]
    # consecutive rank slices tile the stream
    runs = [shard.frames_for_rank(6, 4, r) for r in range(4)]
    edges = [shard.ts_slice_for_frames(hem, a, c) for a, c in runs]
    assert edges[0][0] == 0 and all(edges[i][1] == edges[i + 1][0] for i in range(3)) and edges[-1][1] == pos


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


def _warmup_worker(rank, world, port, q):
    import time
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 1:
        time.sleep(0.3)              # skewed clocks: this rank's timer starts late and its steps are slower
    issued = []

    def step():
        issued.append(len(issued))
        time.sleep(0.002 * (1 + rank))

    def any_rank(flag):
        t = torch.tensor([1 if flag else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(int(t.item()))
    n = shard.collective_warmup(step, lambda: None, any_rank, min_steps=3, min_seconds=0.1)
    q.put((rank, n, len(issued)))
    dist.barrier()
    dist.destroy_process_group()


def test_collective_warmup_issues_the_same_steps_on_every_rank():
    """bench.py's multi-rank warm-up: ranks with skewed clocks and different step times still issue the same number of
    steps (a multiple of the chunk), because the stop decision is an all-reduce -- the property whose absence let a rank
    wait for slot releases the root never issued."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_warmup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got[0][1] == got[1][1] == got[0][2] == got[1][2]
    assert got[0][1] % 8 == 0 and got[0][1] >= 8
