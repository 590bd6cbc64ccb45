"""Whole-step CUDA graph: ``set_input`` + ``update()`` of a ``CSModel`` (reference train.py:212-217) captured once and
replayed per iteration.  One training step is ~4 500 kernel launches of this library; at the reference's own batch size
(4 slices, commands_train_test.sh:32) the step is bound by launch latency, not by the GPU - a replayed
graph removes the per-launch host cost.  Everything inside ``update()`` is device-side (no ``.item()``, no host
branches on tensor values), the AdamW step counter moves to the device (``optim.AdamW.capturable``).

Single-GPU only: the gradient all-reduce of ``parallel.GradBuckets`` is launched from autograd hooks and is not
captured."""
import torch


class GraphedUpdate:
    def __init__(self, net, img_full, img_aux, warmup=3):
        assert net.grad_buckets is None and net.grad_sync is None, "GraphedUpdate: single-process steps only"
        assert img_full.is_cuda and net.training
        self.net = net
        self.img_full, self.img_aux = img_full.clone(), img_aux.clone()          # static input buffers of the graph
        for name in ("optim_G", "optim_D", "optim_T", "optim_R", "optim_M"):
            opt = getattr(net, name, None)
            if opt is not None and hasattr(opt, "capturable"):
                opt.capturable = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # warm-up off the default stream (lazy plans, optimiser state)
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for name in [k for k in net.__dict__ if k.startswith(("loss_", "img_"))]:
            delattr(net, name)                      # drop the warm-up step's tensors: the capture needs their memory
        for opt in (getattr(net, n, None) for n in ("optim_G", "optim_D", "optim_T", "optim_R", "optim_M")):
            if opt is not None:
                opt.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()                    # the graph allocates from its own pool
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
            self.loss_sim = getattr(net, "loss_sim", None)     # static tensors: hold the replay's values
            self.loss_smooth = getattr(net, "loss_smooth", None)
        self.launches_per_step = _lib.launch_count() - n0     # kernels of this library inside one replay
        self._params = [p for n in ("net_T", "net_R", "net_G", "net_D", "net_mask") if getattr(net, n, None) is not None
                        for p in getattr(net, n).parameters()]

    def _step(self):
        self.net.set_input(self.img_full, self.img_aux)
        self.net.update()

    def __call__(self, img_full, img_aux):
        """One training step on a new batch (same shapes): copy into the static buffers, replay."""
        self.img_full.copy_(img_full, non_blocking=True)
        self.img_aux.copy_(img_aux, non_blocking=True)
        self.replay()
        return self.loss_sim

    def replay(self):
        """One step on whatever the static input buffers hold.  The replayed optimiser kernels change the parameters
        without autograd noticing: bump their version counters, so that eager code running afterwards (validation, the
        staged-weight cache of tc.py) does not mistake them for unchanged."""
        self.graph.replay()
        torch.autograd.graph.increment_version(self._params)
