"""Multi-GPU partition of the hot path.

Time slices are independent once the warm start is disabled (reference --stm-disable semantics,
dvs_flow.h:218-219), so the only multi-GPU structure is: cut the stream into slices, give every rank
a block-cyclic share of them, minimise locally (no data-path collective), and gather the fixed-size
per-slice result records once per batch.  One process per GPU; torch.distributed is plumbing only
(NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

RECORD_F64 = 20   # slice id, rc, iters, n_events, flags, 11 model scalars, 4 dividers


def partition(n_slices: int, world: int, rank: int, block: int = 4) -> list[int]:
    """Block-cyclic assignment: blocks of `block` consecutive slices dealt round-robin to the ranks.
    Consecutive slices of a stream have similar iteration counts, so dealing small blocks balances
    the data-dependent GD lengths (26..924 steps in the survey) better than one contiguous chunk."""
    out = []
    for b0 in range(0, n_slices, block):
        if (b0 // block) % world == rank:
            out.extend(range(b0, min(n_slices, b0 + block)))
    return out


def pack_records(slice_ids, results) -> np.ndarray:
    """Result dicts (better_flow_b200.Context.result) -> float64 records for the gather."""
    rec = np.zeros((len(slice_ids), RECORD_F64), dtype=np.float64)
    for k, (sid, r) in enumerate(zip(slice_ids, results)):
        rec[k, 0] = sid
        rec[k, 1] = r["rc"]
        rec[k, 2] = r["iters"]
        rec[k, 3] = r.get("n_events", 0)
        rec[k, 4] = r.get("flags", 0)
        rec[k, 5:16] = r["model"]
        rec[k, 16:20] = r["dividers"]
    return rec


def unpack_record(row) -> dict:
    return {"slice": int(row[0]), "rc": int(row[1]), "iters": int(row[2]), "n_events": int(row[3]),
            "flags": int(row[4]), "model": np.array(row[5:16]), "dividers": np.array(row[16:20], dtype=np.float32)}


def gather_records(local: np.ndarray, n_slices: int, dist=None, device="cpu") -> np.ndarray:
    """All ranks' records in global slice order: ONE all_gather of equal-size buffers (ranks with fewer
    slices pad with slice id -1)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        order = np.argsort(local[:, 0], kind="stable")
        return local[order]
    world = dist.get_world_size()
    cap = max(len(partition(n_slices, world, r)) for r in range(world))
    buf = np.full((cap, RECORD_F64), -1.0)
    buf[:len(local)] = local
    mine = torch.from_numpy(buf).to(device)
    out = torch.empty((world * cap, RECORD_F64), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, mine)
    allrec = out.cpu().numpy()
    allrec = allrec[allrec[:, 0] >= 0]
    return allrec[np.argsort(allrec[:, 0], kind="stable")]


def run_sharded(slices, minimise_batch, dist=None, device="cpu", block: int = 4):
    """Minimise `slices` (the same global list on every rank) across the ranks of `dist`.

    minimise_batch(list_of_slices) -> list of result dicts; on a GPU rank this is a
    better_flow_b200.Context batch run, in the CPU tests it is the oracle.  Returns the records of
    ALL slices in order, on every rank."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    ids = partition(len(slices), world, rank, block)
    results = minimise_batch([slices[i] for i in ids]) if ids else []
    return gather_records(pack_records(ids, results), len(slices), dist, device)


class RecordGather:
    """Deferred gather of the per-slice result records of several batches (bench.py at N > 1).

    Each batch ("step") leaves its rank-local records on the device; `snapshot` copies them into the next slot of a
    device ring (a stream-ordered D2D copy of ~100 KB), `flush` exchanges every slot taken so far with ONE
    all_gather.  A collective per step would make all ranks wait for the slowest one at every step -- and cannot
    overlap the next step either: the persistent minimise kernel holds every register of every SM, so an NCCL
    kernel only runs in the gap between two launches.  Device-agnostic (CUDA + NCCL in bench.py, CPU + gloo in the
    tests)."""

    MAX_SLOTS = 256

    def __init__(self, dist, torch, nbytes: int, slots: int, device):
        self.dist = dist
        self.world = dist.get_world_size()
        self.nbytes = int(nbytes)
        self.slots = max(1, min(int(slots), self.MAX_SLOTS))
        self.ring = torch.empty((self.slots, self.nbytes), dtype=torch.uint8, device=device)
        self.out = torch.empty(self.world * self.slots * self.nbytes, dtype=torch.uint8, device=device)
        self.count = 0
        self.gathers = 0
        self.last = None

    def snapshot(self, records) -> None:
        """records: uint8 tensor of nbytes on the ring's device (this rank's bf_slice_result array of one batch)."""
        if self.count == self.slots:
            self.flush()
        self.ring[self.count].copy_(records.reshape(-1)[:self.nbytes], non_blocking=True)
        self.count += 1

    def flush(self):
        """ONE all_gather of the snapshots taken since the last flush -> tensor [world, batches, nbytes] (or None)."""
        if self.count == 0:
            return None
        n = self.count
        dst = self.out[:self.world * n * self.nbytes]
        self.dist.all_gather_into_tensor(dst, self.ring[:n].reshape(-1))
        self.count = 0
        self.gathers += 1
        self.last = dst.view(self.world, n, self.nbytes)
        return self.last


def deal_pool(local_events, dist, torch, device, block: int = 4):
    """Pooled block-cyclic deal of a weak-scaling batch (bench.py at N > 1).

    Every rank brings the packed 8-byte event records of S slices (``local_events``: list of S numpy arrays of
    better_flow_b200.EVENT_DTYPE).  The POOL is the rank-major concatenation of all ranks' lists -- world * S
    slices -- and it is dealt with `partition` (blocks of `block` consecutive pool slices, round-robin), so every
    rank ends up with S slices drawn evenly from all ranks' streams: iteration counts depend on a stream's contour
    geometry, and a rank that kept its own stream would carry that stream's bias for the whole run (round 1:
    +12 % GD iterations on one rank of eight set the max-over-ranks time).  Set-up plumbing, outside any timed region:
    two all_gathers (slice lengths, padded event bytes), then each rank cuts its share out of the gathered pool on
    `device`.  Returns (global pool ids, list of numpy EVENT_DTYPE arrays) for this rank."""
    from . import EVENT_DTYPE
    world, rank = dist.get_world_size(), dist.get_rank()
    S = len(local_events)
    lens = torch.tensor([len(e) for e in local_events], dtype=torch.int64, device=device)
    all_lens = torch.empty((world, S), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_lens.view(-1), lens)            # (same S on every rank: the collective checks sizes)
    cap = int(all_lens.sum(dim=1).max().item())
    mine = torch.zeros(cap * 8, dtype=torch.uint8, device=device)
    flat = np.concatenate(local_events) if S else np.zeros(0, dtype=EVENT_DTYPE)
    mine[:flat.nbytes] = torch.from_numpy(flat.view(np.uint8).copy()).to(device)
    pool = torch.empty((world, cap * 8), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(pool.view(-1), mine)
    offs = torch.cumsum(all_lens, dim=1) - all_lens                   # first event of every pool slice in its source buffer
    all_lens_h, offs_h = all_lens.cpu().numpy(), offs.cpu().numpy()
    ids = partition(world * S, world, rank, block)
    parts = []
    for g in ids:
        src, i = divmod(g, S)
        o, n = int(offs_h[src, i]), int(all_lens_h[src, i])
        parts.append(pool[src, o * 8:(o + n) * 8])
    got = (torch.cat(parts) if parts else mine[:0]).cpu().numpy().view(EVENT_DTYPE)
    out, k = [], 0
    for g in ids:
        src, i = divmod(g, S)
        n = int(all_lens_h[src, i])
        out.append(got[k:k + n])
        k += n
    return ids, out
