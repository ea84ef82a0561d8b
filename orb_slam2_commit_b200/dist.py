"""Multi-GPU plumbing (one process per GPU, torch.distributed): frame sharding needs no collective; brute-force
matching shards the TRAIN set and exchanges one packed 64-bit word per query.

packed = (dist1 << 48) | (dist2 << 32) | uint32(idx1)   — the layout of orbx_hamming_top2_device.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous shard [start, stop) of n items; contiguity keeps 'first index wins' == lowest global index."""
    per = (n + world - 1) // world
    start = min(n, rank * per)
    return start, min(n, start + per)


def frames_for_rank(n_frames: int, world: int, rank: int):
    """Frames (or stereo pairs) are independent units: frame f -> rank f mod world."""
    return list(range(rank, n_frames, world))


def pack_top2(idx1, dist1, dist2) -> np.ndarray:
    idx = np.asarray(idx1).astype(np.int64) & 0xFFFFFFFF
    return ((np.asarray(dist1).astype(np.uint64) << np.uint64(48)) | (np.asarray(dist2).astype(np.uint64) << np.uint64(32))
            | idx.astype(np.uint64))


def unpack_top2(packed):
    p = np.asarray(packed).astype(np.uint64)
    d1 = (p >> np.uint64(48)).astype(np.int32)
    d2 = ((p >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int32)
    idx = (p & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.int64)
    idx = np.where(d1 >= 256, -1, idx).astype(np.int32)
    return idx, d1, d2


def all_gather_packed(packed, group=None):
    """All-gather one packed word per query from every rank -> tensor [world, nq] (int64 view of the u64 words).
    Works on CUDA tensors over NCCL and on CPU tensors over gloo."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = packed.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view(world, flat.numel())


def hamming_top2_sharded(d_query, d_train_shard, index_base: int, group=None):
    """Config 4: every rank holds a contiguous shard of the train set (global index of its first row = index_base) and
    the full query set. Local top-2 on the GPU, NCCL all-gather of 8 bytes per query per rank, merge kernel.
    Returns CUDA int32 tensors (idx1, dist1, dist2), identical on every rank."""
    import torch
    from . import api
    L = api.lib()
    nq, nt = d_query.shape[0], d_train_shard.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    packed = torch.empty(nq, dtype=torch.int64, device=d_query.device)
    api._ck(L.orbx_hamming_init_device(packed.data_ptr(), nq, st))
    api._ck(L.orbx_hamming_top2_device(d_query.data_ptr(), nq, d_train_shard.data_ptr() if nt else 0, nt, index_base,
                                       packed.data_ptr(), st))
    parts = all_gather_packed(packed, group)
    out = torch.empty((3, nq), dtype=torch.int32, device=d_query.device)
    api._ck(L.orbx_hamming_merge_device(parts.data_ptr(), parts.shape[0], nq, out[0].data_ptr(), out[1].data_ptr(),
                                        out[2].data_ptr(), st))
    return out[0], out[1], out[2]


def hamming_top2_single(d_query, d_train):
    """The same search over a train set that lives on ONE GPU (orbx_hamming_top2_device + the merge kernel with a single
    part): the result the sharded paths must reproduce bit for bit. Returns CUDA int32 tensors (idx1, dist1, dist2)."""
    import torch
    from . import api
    L = api.lib()
    nq, nt = d_query.shape[0], d_train.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    packed = torch.empty(nq, dtype=torch.int64, device=d_query.device)
    api._ck(L.orbx_hamming_init_device(packed.data_ptr(), nq, st))
    api._ck(L.orbx_hamming_top2_device(d_query.data_ptr(), nq, d_train.data_ptr() if nt else 0, nt, 0, packed.data_ptr(), st))
    out = torch.empty((3, nq), dtype=torch.int32, device=d_query.device)
    api._ck(L.orbx_hamming_merge_device(packed.data_ptr(), 1, nq, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), st))
    return out[0], out[1], out[2]


class PeerHammingMatcher:
    """Config 4 with the exchange fused into the matcher kernel (orbx_peer_* in include/orbx.h): the last CTA of every
    query tile stores the rank's top-2 into every peer's landing buffer over NVLink and bumps the peers' arrival
    counters; a one-block kernel waits (bounded) and merges. torch.distributed is used ONCE, at construction, to
    exchange the CUDA-IPC handles; the data path itself makes no NCCL call."""

    def __init__(self, nq_max: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import api
        self.api, self.L = api, api.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.cuda.current_device()
        self.nq_max = nq_max
        self.h = C.c_void_p()
        handle = C.create_string_buffer(64)
        api._ck(self.L.orbx_peer_create(nq_max, self.world, self.rank, self.device, C.byref(self.h), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        api._ck(self.L.orbx_peer_connect(self.h, b"".join(handles)))
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        dist.barrier(group)                      # every rank has mapped every landing buffer before the first store

    def __call__(self, d_query, d_train_shard, index_base: int, check: bool = True):
        """check=True reads the status word back (one 4-byte D2H, synchronises the stream) and raises when a peer did not
        arrive within the kernel's bounded wait; the outputs then hold -1 / 256 ("no match") and the matcher is closed,
        because its arrival counters are out of step with the peers' — build a new one behind a barrier. A caller that
        pipelines calls passes check=False and calls raise_if_failed() once it synchronises anyway."""
        import torch
        if self.h is None:
            raise RuntimeError("PeerHammingMatcher is closed (a previous call timed out or close() was called)")
        nq, nt = d_query.shape[0], d_train_shard.shape[0]
        out = torch.empty((3, nq), dtype=torch.int32, device=d_query.device)
        st = torch.cuda.current_stream().cuda_stream
        self.api._ck(self.L.orbx_peer_hamming_top2(self.h, d_query.data_ptr(), nq, d_train_shard.data_ptr() if nt else 0, nt,
                                                   index_base, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                                   self.status.data_ptr(), st))
        if check:
            self.raise_if_failed()
        return out[0], out[1], out[2]

    def raise_if_failed(self):
        if self.h is not None and int(self.status.item()) != 0:
            self.close()
            raise RuntimeError("PeerHammingMatcher: a peer's top-2 did not arrive within the bounded wait; results are "
                               "'no match' (-1 / 256) and the matcher was closed — recreate it on every rank behind a barrier")

    def close(self):
        if self.h:
            self.L.orbx_peer_destroy(self.h)
            self.h = None
