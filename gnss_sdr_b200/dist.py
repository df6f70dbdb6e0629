"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The two hot paths shard without any data-path exchange except ONE step:
  * tracking: channels are independent given the IQ stream -> round-robin channel ownership;
    every rank that owns a channel on a band receives that band's samples (host fan-out or
    broadcast); taps go back to the owning host thread.  No collective.
  * acquisition: the PRN x Doppler grid is split by PRN; each rank reduces its own sweep to one
    16-byte (statistic, PRN, Doppler bin, code phase) record on the device and the ranks
    all-gather the records; ties resolve like the reference's scans (best_peak).
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


PEAK_DTYPE = np.dtype([("test_statistics", "<f4"), ("prn", "<u4"), ("index_doppler", "<u4"), ("index_time", "<u4")])  # b200_acq_peak


def best_peak(records: np.ndarray) -> np.ndarray:
    """Winner among per-rank b200_acq_peak records: the largest statistic; among equals the lowest PRN, then the lowest
    Doppler bin, then the lowest code phase - the order in which the reference's strict '>' scans keep the first maximum
    (pcps_acquisition.cc:417-426), so a multi-GPU search returns what a single-GPU search returns.  Records with prn == 0
    (a rank that searched nothing) never win."""
    r = np.asarray(records).view(PEAK_DTYPE).reshape(-1)
    valid = r[r["prn"] > 0]
    if valid.size == 0:
        return np.zeros(1, PEAK_DTYPE)[0]
    order = np.lexsort((valid["index_time"], valid["index_doppler"], valid["prn"], -valid["test_statistics"].astype(np.float64)))
    return valid[order[0]]


def allgather_peaks(local_peak):
    """One all-gather of the ranks' 16-byte b200_acq_peak records (torch uint8/int32 tensor of 16 bytes on this rank's
    device, written by b200_acq_sweep_best_dev).  Returns a (world, 4) int32 tensor on the same device - the only
    exchange of the PRN-sharded acquisition.  Unlike a MAX all-reduce over a packed 64-bit key this has
    room for any code-phase index (two-level FFT sizes exceed 2^15) and leaves the tie-breaking to best_peak()."""
    import torch
    import torch.distributed as dist
    t = local_peak.view(torch.int32).reshape(1, 4)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.clone()
    out = torch.empty((dist.get_world_size(), 4), dtype=torch.int32, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out


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
