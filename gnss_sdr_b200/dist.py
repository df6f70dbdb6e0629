"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The two hot paths shard without any data-path exchange except ONE step:
  * tracking: channels are independent given the IQ stream -> round-robin channel ownership;
    every rank that owns a channel on a band receives that band's samples (host fan-out or
    broadcast); taps go back to the owning host thread.  No collective.
  * acquisition: the PRN x Doppler grid is split by PRN; each rank reduces its own rows to
    (peak, index) and the global winner is an all-reduce(MAX) over a packed 64-bit key
    (IEEE-754 bits of a non-negative float are order-preserving as unsigned integers).
The reference has no counterpart (multi-GPU there = cudaSetDevice(rand() % n),
src/algorithms/tracking/libs/cuda_multicorrelator.cu:153-154).
"""
from __future__ import annotations

import numpy as np


def shard_round_robin(n_items: int, world: int, rank: int) -> list:
    """Item ids owned by `rank`: rank, rank + world, ... (channels or PRN slots)."""
    return list(range(rank, n_items, world))


def owner_of(item: int, world: int) -> int:
    return item % world


def pack_peak_key(peak, prn, index_doppler, index_time) -> np.ndarray:
    """(float32 peak >= 0, prn < 256, doppler bin < 256... up to 4095, code phase < 2^20) -> int64 key.
    Layout: [63..32] float bits | [31..24] prn | [23..... split below].
    Bits: peak 32 | prn 8 | doppler bin 9 | index_time 15 = 64."""
    peak = np.asarray(peak, np.float32)
    assert np.all(peak >= 0)
    bits = peak.view(np.uint32).astype(np.uint64)
    prn = np.asarray(prn, np.uint64)
    d = np.asarray(index_doppler, np.uint64)
    t = np.asarray(index_time, np.uint64)
    assert np.all(prn < 256) and np.all(d < 512) and np.all(t < 32768)
    key = (bits << np.uint64(32)) | (prn << np.uint64(24)) | (d << np.uint64(15)) | t
    return key.astype(np.uint64).view(np.int64)   # top bit is the float sign bit = 0, so int64 order == uint64 order


def unpack_peak_key(key):
    k = np.asarray(key, np.int64).view(np.uint64)
    peak = (k >> np.uint64(32)).astype(np.uint32).view(np.float32)
    prn = ((k >> np.uint64(24)) & np.uint64(0xFF)).astype(np.int64)
    d = ((k >> np.uint64(15)) & np.uint64(0x1FF)).astype(np.int64)
    t = (k & np.uint64(0x7FFF)).astype(np.int64)
    return peak, prn, d, t


def allreduce_best_peak(local_key_tensor):
    """In-place MAX all-reduce of an int64 tensor of packed keys (any shape).  One collective."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(local_key_tensor, op=dist.ReduceOp.MAX)
    return local_key_tensor


def gather_results(local_results: np.ndarray, slots_owned: list, n_slots: int, device=None) -> np.ndarray:
    """All ranks end up with the full per-PRN result table (28-byte records), via one all_gather of
    fixed-size padded blocks.  local_results[i] belongs to slots_owned[i]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rec = local_results.dtype.itemsize
    per = (n_slots + world - 1) // world
    buf = np.zeros(per * rec + per * 4, np.uint8)
    buf[: len(slots_owned) * rec] = local_results.view(np.uint8).reshape(-1)[: len(slots_owned) * rec]
    ids = np.full(per, -1, np.int32)
    ids[: len(slots_owned)] = slots_owned
    buf[per * rec:] = ids.view(np.uint8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if world == 1:
        parts = [t]
    else:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
    out = np.zeros(n_slots, local_results.dtype)
    for p in parts:
        b = p.cpu().numpy()
        ids_r = b[per * rec:].view(np.int32)
        recs = b[: per * rec].view(local_results.dtype)
        for k, sid in enumerate(ids_r):
            if sid >= 0:
                out[sid] = recs[k]
    return out


def broadcast_band(block, src: int = 0):
    """Fan one IQ block (torch tensor, complex samples as float32 pairs, resident on this rank's device) out to
    every rank: one broadcast over NCCL/NVLink (8 B * fs per second of signal - 0.4 GB/s for a 50 Msps band,
    SURVEY 8e).  Every rank passes a tensor of the same shape; the source rank's content wins.  Each rank then
    hands its copy to its engine with b200_iq_attach_dev (or keeps a ring of such blocks) and tracks the
    channels shard_round_robin gives it.  Returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(block, src=src)
    return block
