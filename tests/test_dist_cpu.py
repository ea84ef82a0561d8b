"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU paths — contiguous train
shards, frame -> rank assignment, the packed (dist1, dist2, idx1) word and its all-gather — checked against the
single-shard oracle. The per-shard top-2 itself comes from the oracle here (no GPU); on the GPU box the same
plumbing carries the CUDA kernel's output (tests/test_gpu_parity.py::test_hamming_shards_merge_on_device)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam2_commit_b200 import dist as od
from orb_slam2_commit_b200 import synth


def merge_numpy(parts):
    """The merge rule of SURVEY.md §8e, restated for the test: best = min d1 (lowest index on ties),
    second = 2nd smallest of the multiset of all d1 and d2."""
    idx, d1, d2 = od.unpack_top2(parts[0])
    idx = idx.astype(np.int64); idx[idx < 0] = 0xFFFFFFFF
    for p in parts[1:]:
        i2, e1, e2 = od.unpack_top2(p)
        i2 = i2.astype(np.int64); i2[i2 < 0] = 0xFFFFFFFF
        first = (d1 < e1) | ((d1 == e1) & (idx <= i2))
        d2 = np.minimum(np.maximum(d1, e1), np.minimum(d2, e2))
        idx = np.where(first, idx, i2); d1 = np.where(first, d1, e1)
    idx = np.where(d1 >= 256, -1, idx).astype(np.int32)
    return idx, d1, d2


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, False, repr(e), None))
        raise


def _worker_body(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    train, query = synth.synth_descriptors(6000, 96, seed=7)
    query[3] = ~train[10]
    a, b = od.shard_range(len(train), world, rank)
    i1, d1, d2 = ob.hamming_top2(query, train[a:b])
    gi = np.where(i1 >= 0, i1 + a, -1)
    packed = torch.from_numpy(od.pack_top2(gi, d1, d2).view(np.int64))
    parts = od.all_gather_packed(packed).numpy().view(np.uint64)
    got = merge_numpy(list(parts))
    want = ob.hamming_top2(query, train)
    ok = all(np.array_equal(g, w) for g, w in zip(got, want))
    frames = od.frames_for_rank(11, world, rank)
    q.put((rank, ok, frames, (a, b)))
    dist.barrier()
    dist.destroy_process_group()


def test_train_sharding_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] and res[1][1], "merged per-shard top-2 differs from the single-shard result"
    assert sorted(res[0][2] + res[1][2]) == list(range(11))          # every frame extracted exactly once
    assert res[0][3] == (0, 3000) and res[1][3] == (3000, 6000)


def test_pack_unpack_roundtrip_and_ranges():
    rng = np.random.default_rng(0)
    idx = rng.integers(-1, 1 << 31, 1000).astype(np.int64); d1 = rng.integers(0, 257, 1000); d2 = np.maximum(d1, rng.integers(0, 257, 1000))
    idx[d1 >= 256] = -1
    i, a, b = od.unpack_top2(od.pack_top2(idx, d1, d2))
    assert np.array_equal(i, idx.astype(np.int32)) and np.array_equal(a, d1) and np.array_equal(b, d2)
    for n in (0, 1, 7, 1_000_000):
        for world in (1, 2, 4, 8):
            r = [od.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
