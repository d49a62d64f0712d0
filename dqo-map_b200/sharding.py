"""Object-sharded mapping across GPUs.

The reference optimises one scene-level cloud on one GPU and has no distributed code.  What shards naturally is
the object dimension: every Gaussian carries an object id (`_obj_id`, SLAM/gaussian_pointcloud.py:497,54) and every
object owns an independently optimised quadric (quadrics.py:2245-2295).  Rank r maps the objects assigned to it with
no gradient exchange; NCCL (gloo in the CPU tests) is used only to gather the small per-object table
(pose / quadric parameters) and, on demand, the Gaussians themselves to rank 0 for tracking / whole-scene rendering.
"""
import torch
import torch.distributed as dist


def assign_objects(counts, world_size):
    """Greedy longest-processing-time bin packing of objects by Gaussian count.
    counts: dict {obj_id: n_gaussians} or sequence indexed by object id.  Returns (owner dict, per-rank load list).
    Deterministic: ties broken by object id, so every rank computes the same assignment without communication."""
    items = sorted(counts.items() if isinstance(counts, dict) else enumerate(counts), key=lambda kv: (-int(kv[1]), kv[0]))
    load = [0] * world_size
    owner = {}
    for oid, c in items:
        r = min(range(world_size), key=lambda i: (load[i], i))
        owner[oid] = r
        load[r] += int(c)
    return owner, load


def local_objects(owner, rank):
    return sorted(o for o, r in owner.items() if r == rank)


def gather_object_table(local_table, group=None, rows_per_rank=None):
    """all_gather of a per-object float table [n_local, F] with variable n_local (e.g. F = 12:
    obj_id, count, center 3, quaternion 4, axes 3).  Returns the concatenated [n_total, F] table on every rank,
    sorted by column 0 (object id).  When every rank is known to hold exactly `rows_per_rank` rows (static object
    assignment; ranks with fewer objects pad with rows whose id is negative) the size exchange and its host
    synchronisation are skipped: one collective, fully asynchronous, same ordering."""
    world = dist.get_world_size(group)
    dev = local_table.device
    if rows_per_rank is not None:
        if local_table.shape[0] != rows_per_rank:
            raise ValueError("rows_per_rank does not match the local table")
        out = torch.empty((world * rows_per_rank, local_table.shape[1]), dtype=local_table.dtype, device=dev)
        dist.all_gather_into_tensor(out, local_table.contiguous(), group=group)
        # same row order as the general path: by object id (device-side argsort, no host synchronisation); padding rows of
        # ranks that own fewer objects (id < 0) come first
        return out[torch.argsort(out[:, 0], stable=True)] if out.shape[0] > 1 else out
    n_local = torch.tensor([local_table.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    F = local_table.shape[1]
    mx = max(max(sizes), 1)
    padded = torch.zeros((mx, F), dtype=local_table.dtype, device=dev)
    padded[: local_table.shape[0]] = local_table
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    table = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    if table.shape[0] > 1:
        table = table[torch.argsort(table[:, 0], stable=True)]
    return table


def gather_gaussians(tensors, dst=0, group=None):
    """Gathers variable-length per-rank Gaussian tensors (dict name -> [n_local, ...]) to rank `dst`.
    Returns the concatenated dict on `dst`, None elsewhere.  Counts travel first, payloads as padded gathers."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    names = sorted(tensors)
    any_t = tensors[names[0]]
    dev = any_t.device
    n_local = torch.tensor([any_t.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    out = {}
    for name in names:
        t = tensors[name].contiguous()
        padded = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        padded[: t.shape[0]] = t
        bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, bufs, dst=dst, group=group)
        if rank == dst:
            out[name] = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    return out if rank == dst else None
