"""Multi-GPU plumbing: one process per GPU, utterances sharded across ranks, weights moved with
ONE broadcast of a flat fp32 buffer; no collective on the hot path.

The reference has no inference parallelism (reconstruct.py:17 is single-device); training uses
DDP whose only inference-relevant action is the initial parameter/buffer broadcast from rank 0
(train.py:49-50 broadcast_buffers=True).  This module is that broadcast plus the shard map.
Works with backend 'nccl' (GPU tensors, NVLink) and 'gloo' (CPU tensors; used by the tests)."""
from collections import OrderedDict

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block partition of n_items utterances: first (n % world) ranks get one extra."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def flatten_state_dicts(sds):
    """{'encoder': sd, ...} -> (flat fp32 tensor, layout) ; layout = [(module, key, shape, offset, numel)]."""
    layout, chunks, off = [], [], 0
    for mod in sorted(sds.keys()):
        for k, v in sds[mod].items():
            n = v.numel()
            layout.append((mod, k, tuple(v.shape), off, n))
            chunks.append(v.detach().reshape(-1).to(torch.float32).cpu())
            off += n
    return torch.cat(chunks) if chunks else torch.zeros(0), layout


def unflatten_state_dicts(flat, layout):
    out = {}
    flat = flat.cpu()
    for mod, k, shape, off, n in layout:
        out.setdefault(mod, OrderedDict())[k] = flat[off:off + n].reshape(shape).clone()
    return out


def broadcast_state_dicts(sds, src=0, device=None):
    """Rank `src` passes the checkpoint state_dicts (others pass None); every rank returns the same
    state_dicts.  One object broadcast of the (tiny) layout, then ONE tensor broadcast of all weights."""
    rank = dist.get_rank()
    if rank == src:
        flat, layout = flatten_state_dicts(sds)
    else:
        flat, layout = None, None
    box = [layout]
    dist.broadcast_object_list(box, src=src)
    layout = box[0]
    total = sum(n for *_, n in layout)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = flat.to(device) if rank == src else torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    return unflatten_state_dicts(buf, layout)


def max_over_ranks(value, device=None):
    """Device-timed milliseconds etc.: MAX over ranks (bench.py contract)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
