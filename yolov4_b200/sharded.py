"""Image-sharded multi-GPU form of the hot path.

Images are independent in decode, filter, sort and NMS (utils.py:133 loops per image), so rank r of G simply owns the
contiguous image range shard_range(B, r, G) and runs the single-GPU path on it: no collective on the hot path.
Only the final detections are exchanged: one all-gather of the per-image counts and one of the rows, trimmed to
the largest per-image count so no padding beyond that travels over NVLink.
"""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous, balanced image range [lo, hi) of rank `rank` (first n_images % world ranks get one extra)."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_detections(rows, counts, group=None):
    """rows [B_local, cap, 7] (padded, device or CPU), counts [B_local] int32 -> list over ALL images (rank-major) of
    [K_i, 7] tensors or None, identical on every rank.  All ranks must hold the same B_local (pad the last shard)."""
    world = dist.get_world_size(group)
    b_local = counts.shape[0]
    all_counts = torch.empty((world * b_local,), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
    counts_host = all_counts.cpu()
    k_max = int(counts_host.max()) if counts_host.numel() else 0
    if k_max > rows.shape[1]:
        raise RuntimeError("a rank truncated its rows (count %d > capacity %d)" % (k_max, rows.shape[1]))
    k_max = max(k_max, 1)
    send = rows[:, :k_max].contiguous()
    recv = torch.empty((world * b_local, k_max, 7), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = []
    for i in range(world * b_local):
        k = int(counts_host[i])
        out.append(recv[i, :k] if k else None)
    return out
