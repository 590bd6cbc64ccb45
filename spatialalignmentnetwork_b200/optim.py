"""``AdamW`` with the arithmetic of ``torch.optim.AdamW`` (no amsgrad; reference model.py:72-81 builds one per
network with lr 1e-4, weight_decay 0) stepping ALL tensors of a parameter group with one multi-tensor launch
per 48 tensors of ``san_adamw_step``.  State layout (``step`` / ``exp_avg`` /
``exp_avg_sq`` per parameter) and ``state_dict`` format are torch's, so optimiser checkpoints interoperate."""
import ctypes

import torch

from ._lib import call


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        assert lr >= 0 and eps >= 0 and 0 <= betas[0] < 1 and 0 <= betas[1] < 1 and weight_decay >= 0
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # capturable = True (graphs.GraphedUpdate sets it): the step counter of the bias corrections lives on the device,
        # so that a captured CUDA graph advances it on every replay (the analogue of torch's ``capturable=True``)
        self.capturable = False
        self._dev_step = {}

    def _device_step(self, group_index, device, step):
        if group_index not in self._dev_step:
            self._dev_step[group_index] = (torch.full((1,), step, dtype=torch.int32, device=device),
                                           torch.zeros(2, dtype=torch.float32, device=device))
        return self._dev_step[group_index]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            buckets = {}            # step count -> ([p], [g], [m], [v])
            for p in group["params"]:
                if p.grad is None:
                    continue
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous(), \
                    "san_b200 AdamW: contiguous fp32 CUDA parameters (no CPU fallback)"
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                b = buckets.setdefault(st["step"], ([], [], [], []))
                b[0].append(p); b[1].append(g); b[2].append(st["exp_avg"]); b[3].append(st["exp_avg_sq"])
            if self.capturable:
                assert len(buckets) == 1, "capturable AdamW: all parameters of a group share one step count"
            for step, (ps, gs, ms, vs) in buckets.items():
                n = len(ps)
                ptrs = [(ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]) for ts in (ps, gs, ms, vs)]
                numel = (ctypes.c_longlong * n)(*[t.numel() for t in ps])
                step_dev, sched_dev = self._device_step(gi, ps[0].device, step - 1) if self.capturable else (None, None)
                call("adamw_step", *[ctypes.addressof(a) for a in ptrs], ctypes.addressof(numel), n,
                     float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                     float(group["weight_decay"]), step, step_dev, sched_dev)
                # the kernel wrote through raw pointers: tell autograd (and the staged-weight cache of tc.py, which keys on
                # the version counter) that the parameters changed
                torch.autograd.graph.increment_version(ps)
        return loss
