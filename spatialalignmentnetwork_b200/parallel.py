"""One-process-per-GPU data parallelism for the training step (all ``reg`` modes): slices are independent, so the batch
is sharded across ranks and the only exchange is a gradient all-reduce (mean) per optimiser
step over NCCL / NVLink (gloo on CPU for tests).  The reference has no distributed code
(SURVEY.md §2.2); BatchNorm in ``net_T`` uses per-rank batch statistics (standard DDP
semantics), identical masks are obtained by seeding python ``random`` identically."""
import torch
import torch.distributed as dist


def flatten_grads(params):
    params = [p for p in params if p.grad is not None]
    if not params:
        return params, None
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    return params, flat


def unflatten_grads(params, flat):
    o = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n


def allreduce_mean_grads(params, group=None):
    """Average gradients over ranks with ONE all-reduce on a flat fp32 bucket."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params, flat = flatten_grads(list(params))
    if flat is None:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    unflatten_grads(params, flat)


def broadcast_state(modules, src=0, group=None):
    """Make parameters and buffers identical on all ranks (rank ``src`` wins)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def attach(model, group=None):
    """Install the gradient all-reduce into a ``CSModel`` and synchronise its initial state."""
    nets = [getattr(model, k) for k in ("net_mask", "net_G", "net_D", "net_T", "net_R") if hasattr(model, k)]
    broadcast_state(nets, group=group)
    model.grad_sync = lambda params: allreduce_mean_grads(params, group=group)
    return model


def shard(batch, rank, world_size):
    """Rank r gets slices [r*N/G, (r+1)*N/G)."""
    n = batch.shape[0]
    assert n % world_size == 0, "global batch must divide evenly over ranks"
    per = n // world_size
    return batch[rank * per:(rank + 1) * per]
