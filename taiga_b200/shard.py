"""Multi-GPU sharding of a batch of partial transactions.

Proofs are independent, so a batch of B ptx is split into contiguous blocks, one per rank (ptx i -> rank floor(i*G/B));
every rank holds a full replica of the SRS and both proving keys and proves its block with the single-GPU batched
engine.  The only collective on the path is the gather of the finished proof bytes (fixed-size records), done with
`torch.distributed.all_gather` - NCCL over NVLink on the GPUs, gloo in the CPU tests.  Nothing inside a proof crosses
ranks.  (The reference builds the proofs of `ShieldedPartialTransaction::build` one after the other on one host,
shielded_ptx.rs:107-125.)
"""
import numpy as np


def shard_range(n_items, world, rank):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def pack_records(compliance_proofs, vp_proofs, len_c, len_v, per_ptx=(2, 4)):
    """One fixed-size record per ptx: 2 Compliance proofs followed by 4 VP proofs."""
    nc, nv = per_ptx
    n_ptx = len(compliance_proofs) // nc
    assert len(compliance_proofs) == nc * n_ptx and len(vp_proofs) == nv * n_ptx
    rec = np.zeros((n_ptx, nc * len_c + nv * len_v), np.uint8)
    for i in range(n_ptx):
        parts = compliance_proofs[nc * i: nc * (i + 1)] + vp_proofs[nv * i: nv * (i + 1)]
        rec[i] = np.frombuffer(b"".join(parts), np.uint8)
    return rec


def unpack_record(rec, len_c, len_v, per_ptx=(2, 4)):
    nc, nv = per_ptx
    b = bytes(rec)
    cp = [b[i * len_c:(i + 1) * len_c] for i in range(nc)]
    off = nc * len_c
    vp = [b[off + i * len_v: off + (i + 1) * len_v] for i in range(nv)]
    return cp, vp


def gather_records(local_records, n_total, device=None):
    """all_gather of the per-rank record blocks into the full [n_total, record_len] array (every rank gets it).
    Blocks may differ in length by one ptx, so they are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    rec_len = local_records.shape[1]
    pad = np.zeros((max_rows, rec_len), np.uint8)
    pad[: local_records.shape[0]] = local_records
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t)
    out = np.zeros((n_total, rec_len), np.uint8)
    for r, (lo, hi) in enumerate(sizes):
        out[lo:hi] = bufs[r][: hi - lo].cpu().numpy()
    return out
