"""Clip-level data parallelism: clips and trajectories are independent, so every rank runs
the whole path on its own shard; the only exchange is one gather of fixed-size result records per run.

Record layout per clip (616 bytes): [T' (stored as float32), spin[3], pos[50][3]] float32."""
import torch
import torch.distributed as dist

SEQ_LEN = 50
RECORD_FLOATS = 1 + 3 + SEQ_LEN * 3


def shard_range(n_items, rank, world):
    """Contiguous block partition; the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_record(spin, pos3d):
    """spin (3,), pos3d (T', 3) -> (RECORD_FLOATS,) float32 tensor on spin's device."""
    rec = torch.zeros(RECORD_FLOATS, dtype=torch.float32, device=spin.device)
    n = min(int(pos3d.shape[0]), SEQ_LEN)
    rec[0] = n
    rec[1:4] = spin.to(torch.float32)
    rec[4:4 + 3 * n] = torch.as_tensor(pos3d[:n], dtype=torch.float32, device=spin.device).reshape(-1)
    return rec


def unpack_record(rec):
    n = int(rec[0].item())
    return rec[1:4], rec[4:4 + 3 * n].reshape(n, 3)


def gather_records(local, n_total, world=None):
    """local: (n_local, RECORD_FLOATS) records of this rank's block (shard_range order).
    Returns the (n_total, RECORD_FLOATS) table on every rank.  NCCL on GPUs, gloo on CPU."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size() if world is None else world
    per = (n_total + world - 1) // world
    padded = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    rows = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        rows.append(out[r * per:r * per + (b - a)])
    return torch.cat(rows)


def run_clips(predict_clip, clips, device):
    """Shard `clips` over the ranks, run predict_clip(clip) -> (spin, pos3d) on this rank's block, gather all records."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    a, b = shard_range(len(clips), rank, world)
    recs = [pack_record(*predict_clip(clips[i])) for i in range(a, b)]
    local = torch.stack(recs) if recs else torch.zeros((0, RECORD_FLOATS), dtype=torch.float32, device=device)
    return gather_records(local.to(device), len(clips), world)
