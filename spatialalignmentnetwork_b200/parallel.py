"""One-process-per-GPU data parallelism for the training step (all ``reg`` modes): slices are independent, so the batch
is sharded across ranks and the only exchange is the gradient all-reduce (mean) of each optimiser step over NCCL /
NVLink (gloo on CPU for tests).  The reference has no distributed code.

Gradient exchange (SURVEY.md §8e "bucketed by net and by cascade, overlapped"): every network owns ONE flat fp32
gradient buffer; each parameter's ``.grad`` is a view into it, so there is no flatten / copy-back pass.  The buffer is
cut into buckets (one per VarNet cascade, one for the sensitivity net, one per other network); a
``post_accumulate_grad`` hook counts the bucket's parameters down during ``backward()`` and launches the bucket's
all-reduce asynchronously the moment its last gradient is written - cascade 11's gradients are complete first and
travel while cascades 10..0 are still being differentiated.  ``CSModel.update`` brackets each backward with
``arm(nets)`` / ``sync(nets)``; ``sync`` only waits for the handles.

BatchNorm in ``net_T`` / ``net_G`` uses per-rank batch statistics (standard DDP semantics) unless
``attach(..., sync_bn=True)`` asks for global-batch statistics (the per-plane sums of all ranks are all-gathered before
the finalise kernels, ``tc.SYNC_BN_GROUP``); identical masks are obtained by seeding python ``random`` identically."""
import torch
import torch.distributed as dist

_NETS = ("net_mask", "net_G", "net_D", "net_T", "net_R")


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


# ------------------------------------------------------------------------------------------- simple (blocking) path
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
    """Average gradients over ranks with ONE blocking all-reduce on a flat fp32 bucket (the fallback for parameter
    lists that are not managed by ``GradBuckets``)."""
    if not _active(group):
        return
    params, flat = flatten_grads(list(params))
    if flat is None:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    unflatten_grads(params, flat)


# ------------------------------------------------------------------------------------------- bucketed, overlapped path
def _bucket_key(net_name, pname):
    """Bucket of a parameter: per cascade for the VarNet (reference varnet.py:443-445 ``cascades.<i>.``), else per net."""
    if net_name == "net_R" and pname.startswith("cascades."):
        return "cascades." + pname.split(".")[1]
    if net_name == "net_R":
        return pname.split(".")[0]
    return net_name


class _Bucket:
    __slots__ = ("net", "key", "lo", "hi", "nparams", "pending", "handle")

    def __init__(self, net, key, lo):
        self.net, self.key, self.lo, self.hi, self.nparams, self.pending, self.handle = net, key, lo, lo, 0, 0, None


class GradBuckets:
    """Flat per-network gradient buffers whose buckets are all-reduced from autograd hooks while the backward runs."""

    def __init__(self, model, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.avg_op = dist.ReduceOp.AVG if dist.get_backend(group) == "nccl" else None   # gloo has no AVG
        self.flat, self.views, self.buckets, self.params = {}, {}, {}, {}
        self.armed = set()
        self.launched = 0          # all-reduces launched from hooks (i.e. overlapped with the backward), for tests / bench
        for name in _NETS:
            net = getattr(model, name, None)
            if net is None:
                continue
            named = [(k, p) for k, p in net.named_parameters() if p.requires_grad]
            if not named:
                continue
            assert all(p.dtype == torch.float32 for _, p in named)
            total = sum(p.numel() for _, p in named)
            flat = torch.zeros(total, dtype=torch.float32, device=named[0][1].device)
            views, buckets, o = [], [], 0
            for k, p in named:
                key = _bucket_key(name, k)
                if not buckets or buckets[-1].key != key:
                    buckets.append(_Bucket(name, key, o))
                b = buckets[-1]
                views.append(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
                b.hi, b.nparams = o, b.nparams + 1
                p.register_post_accumulate_grad_hook(self._make_hook(b))
            self.flat[name], self.views[name], self.buckets[name], self.params[name] = flat, views, buckets, [p for _, p in named]

    def _make_hook(self, bucket):
        def hook(_param):
            if bucket.net in self.armed and bucket.handle is None:
                bucket.pending -= 1
                if bucket.pending == 0:
                    self._launch(bucket)
                    self.launched += 1
        return hook

    def _launch(self, b):
        seg = self.flat[b.net][b.lo:b.hi]
        b.handle = dist.all_reduce(seg, op=self.avg_op or dist.ReduceOp.SUM, group=self.group, async_op=True)

    def arm(self, names):
        """Call right before ``backward()`` (after ``zero_grad``): gradients of the named networks accumulate into
        their zeroed flat buffers; complete buckets leave immediately."""
        for name in names:
            if name not in self.flat:
                continue
            self.flat[name].zero_()
            for p, v in zip(self.params[name], self.views[name]):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v
            for b in self.buckets[name]:
                b.pending, b.handle = b.nparams, None
            self.armed.add(name)

    def sync(self, names):
        """Call after ``backward()``: launch what the hooks did not (buckets with parameters that received no
        gradient), wait for every bucket, finish the mean."""
        for name in names:
            if name not in self.armed:
                continue
            for b in self.buckets[name]:
                if b.handle is None:
                    self._launch(b)
            for b in self.buckets[name]:
                b.handle.wait()
                b.handle = None
            if self.avg_op is None:
                self.flat[name].div_(self.world)
            self.armed.discard(name)


def broadcast_state(modules, src=0, group=None):
    """Make parameters and buffers identical on all ranks (rank ``src`` wins)."""
    if not _active(group):
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)
            torch.autograd.graph.increment_version(t)     # written through .data: caches keyed on the version must see it


def average_buffers(model, group=None):
    """Average the floating-point buffers (BatchNorm running statistics, spectral-norm vectors) of every network
    over the ranks: they evolve per rank during training (per-rank batch statistics), so eval-mode outputs and
    saved checkpoints would otherwise differ between ranks."""
    if not _active(group):
        return
    w = dist.get_world_size(group)
    for k in _NETS:
        net = getattr(model, k, None)
        if net is None:
            continue
        bufs = [b for b in net.buffers() if b.is_floating_point()]
        if not bufs:
            continue
        flat = torch.cat([b.reshape(-1) for b in bufs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(w)
        o = 0
        for b in bufs:
            b.copy_(flat[o:o + b.numel()].view_as(b))
            o += b.numel()


def attach(model, group=None, overlap=True, sync_bn=False):
    """Install the gradient exchange into a ``CSModel`` (after ``model.to(device)``) and synchronise its initial
    state.  ``overlap=False`` keeps the blocking single-bucket all-reduce after the backward; ``sync_bn=True`` makes
    every BatchNorm layer (``net_T``, ``net_G``, ``net_D``) use the statistics of the GLOBAL batch (tc.SYNC_BN_GROUP)."""
    nets = [getattr(model, k) for k in _NETS if hasattr(model, k)]
    broadcast_state(nets, group=group)
    if not _active(group):
        return model
    if sync_bn:
        from . import tc
        tc.SYNC_BN_GROUP = group if group is not None else dist.group.WORLD
    if overlap:
        model.grad_buckets = GradBuckets(model, group=group)
    model.grad_sync = lambda params: allreduce_mean_grads(params, group=group)
    return model


def shard(batch, rank, world_size):
    """Rank r gets slices [r*N/G, (r+1)*N/G)."""
    n = batch.shape[0]
    assert n % world_size == 0, "global batch must divide evenly over ranks"
    per = n // world_size
    return batch[rank * per:(rank + 1) * per]
