"""Two-GPU tests (skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
data-parallel semantics of the alignment network under ``parallel.attach(..., sync_bn=True)``: with the BatchNorm
statistics of the GLOBAL batch a step sharded over two ranks equals the single-process step at the global batch size
(the reference is single-process: unet.py:119-140 BatchNorm2d in training mode reduces over the whole batch)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(sys.argv[3])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2,
                        device_id=torch.device("cuda", rank))
from spatialalignmentnetwork_b200 import parallel, tc
from spatialalignmentnetwork_b200.cross import SpatialTransformer

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

class Model:
    pass

torch.manual_seed(11)
net = SpatialTransformer(channels=1)
with torch.no_grad():
    torch.nn.init.normal_(net.net[-1].weight, 0, 1e-2)
sd = {k: v.clone() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(12)
moving, fixed = torch.rand(4, 1, 48, 64, generator=g), torch.rand(4, 1, 48, 64, generator=g)

def step(n, mv, fx):
    n.train()
    for p in n.parameters():
        p.grad = None
    offset, grid = n(moving=mv, fixed=fx)
    warped = n.warp(mv, grid)
    loss = offset.pow(2).mean() * 10 + (warped - fx).pow(2).mean()      # means over the (local) batch
    loss.backward()
    return offset.detach(), {k: p.grad.detach().clone() for k, p in n.named_parameters()}

# single-process reference at the global batch (both ranks compute it, on their own GPU)
ref = SpatialTransformer(channels=1).cuda()
ref.load_state_dict(sd)
off_ref, g_ref = step(ref, moving.cuda(), fixed.cuda())
bn_ref = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k}

# sharded over two ranks with global-batch BatchNorm statistics
m = Model()
m.net_T = SpatialTransformer(channels=1).cuda()
m.net_T.load_state_dict(sd)
parallel.attach(m, overlap=False, sync_bn=True)
assert tc.SYNC_BN_GROUP is not None
sl = slice(2 * rank, 2 * rank + 2)
off, grads = step(m.net_T, moving[sl].cuda(), fixed[sl].cuda())
parallel.allreduce_mean_grads(m.net_T.parameters())
grads = {k: p.grad.detach().clone() for k, p in m.net_T.named_parameters()}
assert rel(off, off_ref[sl]) < 2e-5, rel(off, off_ref[sl])
worst = max((rel(grads[k], g_ref[k]), k) for k in g_ref if g_ref[k].norm() > 1e-6 * max(v.norm() for v in g_ref.values()))
assert worst[0] < 2e-3, worst
for k, v in m.net_T.state_dict().items():
    if "running" in k:
        assert rel(v, bn_ref[k]) < 1e-5, k

# without sync_bn the per-rank statistics give a different (standard DDP) result: the test above is not vacuous
tc.SYNC_BN_GROUP = None
m2 = SpatialTransformer(channels=1).cuda()
m2.load_state_dict(sd)
off2, _ = step(m2, moving[sl].cuda(), fixed[sl].cuda())
assert rel(off2, off_ref[sl]) > 1e-4
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sync_bn_matches_global_batch(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29700 + os.getpid() % 50)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[-3000:] for o in outs]
