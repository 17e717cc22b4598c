"""One-process-per-GPU plumbing for the batch-sharded inference path.

Inference has NO data-path collective: every rank owns a contiguous shard of the questions and a
full weight replica.  torch.distributed is used only to agree on timings (max over ranks) and, in
tests, to gather the per-shard answers.  Backend: "nccl" on GPUs, "gloo" in the CPU tests.
"""
import os

import torch
import torch.distributed as dist

from . import synth


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for world 1)."""
    rank, world, _ = env_rank_world()
    if world <= 1:
        return rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        kw = {}
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (device timings -> job time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def shard_for_rank(batch, rank=None, world=None):
    """This rank's contiguous slice of the questions and of their item rows."""
    if rank is None or world is None:
        r, w, _ = env_rank_world()
        rank = r if rank is None else rank
        world = w if world is None else world
    return synth.shard_batch(batch, rank, world)


def gather_picks(picks):
    """All ranks' answer indices, concatenated in rank order (tests / reporting only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(picks)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(picks))
    return [p for part in out for p in part]


def allreduce_mean_grads(parameters):
    """The ONE collective of the (optional) data-parallel training step: the
    trainable parameters' gradients are flattened into a single fp32 bucket (12.25 M elements,
    49 MB with the shipped conf), all-reduced (sum) over NCCL / NVLink and divided by the world
    size, between `backward()` and `clip_grad_norm_` (SDNetTrainer.py:362-366).  One bucket: at
    NVSwitch bandwidth the transfer is ~0.1 ms, so there is nothing to overlap.
    Returns the number of elements reduced.  (The forward kernels of this build have no backward;
    the helper is the plumbing a training path would call.)"""
    grads = [p.grad for p in parameters if p.requires_grad and p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off
