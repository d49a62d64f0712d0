"""Object-sharded mapping across GPUs.

The reference optimises one scene-level cloud on one GPU and has no distributed code.  What shards naturally is
the object dimension: every Gaussian carries an object id (`_obj_id`, SLAM/gaussian_pointcloud.py:497,54) and every
object owns an independently optimised quadric (quadrics.py:2245-2295).  Rank r maps the objects assigned to it with
no gradient exchange; NCCL (gloo in the CPU tests) is used only to gather the small per-object table
(pose / quadric parameters) and, on demand, the Gaussians themselves to rank 0 for tracking / whole-scene rendering.
"""
import math

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


def gather_gaussians(tensors, dst=0, group=None, sizes=None):
    """Gathers variable-length per-rank Gaussian tensors (dict name -> [n_local, ...] float32) to rank `dst`.
    Returns the concatenated dict on `dst`, None elsewhere.

    Every rank packs its tensors into ONE [n_local, F] buffer (F = 59 floats for the six parameter groups of the
    mapping step) and sends exactly its own rows; `dst` posts one receive per peer straight into the row range of the
    result, so nothing is padded, nothing is re-concatenated and there is one message per peer instead of one
    collective per tensor.  `sizes` (rows per rank): pass it when the assignment is static (sharding.assign_objects makes
    it known everywhere) and the size exchange with its host synchronisation is skipped."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    names = sorted(tensors)
    any_t = tensors[names[0]]
    dev = any_t.device
    n_here = any_t.shape[0]
    if sizes is None:
        n_local = torch.tensor([n_here], dtype=torch.int64, device=dev)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        sizes = [int(s.item()) for s in got]
    if sizes[rank] != n_here:
        raise ValueError("sizes[rank] does not match the local tensors")
    widths = [math.prod(tensors[n].shape[1:]) for n in names]
    packed = torch.cat([tensors[n].reshape(n_here, -1).to(torch.float32) for n in names], dim=1).contiguous()
    F = packed.shape[1]
    if rank != dst:
        if n_here:  # batched like the receives on `dst` (an unbatched send is serialised with every other op of the group)
            peer = dist.get_global_rank(group, dst) if group is not None else dst
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, packed, peer, group)]):
                req.wait()
        return None
    total = sum(sizes)
    out = torch.empty((total, F), dtype=torch.float32, device=dev)
    off, ops = 0, []
    for r in range(world):
        rows = out[off:off + sizes[r]]
        if r == dst:
            rows.copy_(packed)
        elif sizes[r]:
            src = dist.get_global_rank(group, r) if group is not None else r
            ops.append(dist.P2POp(dist.irecv, rows, src, group))
        off += sizes[r]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    result, col = {}, 0
    for n, w in zip(names, widths):
        result[n] = out[:, col:col + w].reshape((total,) + tuple(tensors[n].shape[1:])).contiguous()
        col += w
    return result
