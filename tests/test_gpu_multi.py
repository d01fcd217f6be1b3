"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU under torchrun,
NCCL.  The sharded computation must reproduce the single-GPU one on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CP_ROOT"])
from contrastiveprosthetics_b200 import clip as C, dist as cpdist
rank, world, dev = cpdist.init_from_env()
assert dev.type == "cuda" and world >= 2

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

# ---- batch x batch CLIP head: sharded (all-gather / all-reduce / reduce-scatter) == single GPU on the whole batch
n = 1000
g = torch.Generator().manual_seed(0)
E_all = torch.randn(world * n, 16, generator=g).to(dev)
G_all = (torch.randn(world * n, 16, generator=g) * 2.0).to(dev) + 0.7 * E_all
for logit_scale in (0.0, 1.0):
    E = E_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    G = G_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    loss, ncor, arg = C.clip_head(E, G, logit_scale)
    loss.backward()
    Ef = E_all.clone().requires_grad_(True)
    Gf = G_all.clone().requires_grad_(True)
    loss_f, ncor_f, arg_f = C.clip_head(Ef, Gf, logit_scale, group="local")
    loss_f.backward()
    assert abs(loss.item() - loss_f.item()) < 2e-6 * abs(loss_f.item()), (loss.item(), loss_f.item())
    assert rel(E.grad, Ef.grad[rank * n:(rank + 1) * n]) < 1e-5
    assert rel(G.grad, Gf.grad[rank * n:(rank + 1) * n]) < 1e-5
    assert int(ncor) == int(ncor_f) and torch.equal(arg, arg_f[rank * n:(rank + 1) * n])

# ---- flat-bucket gradient all-reduce over NCCL: average and sum
ps = [torch.nn.Parameter(torch.zeros(1000, 3, device=dev)), torch.nn.Parameter(torch.zeros(7, device=dev))]
for mode, expect in ((True, (world + 1) / 2.0), (False, world * (world + 1) / 2.0)):
    for p in ps:
        p.grad = torch.full_like(p, float(rank + 1))
    cpdist.FlatGradAllReduce(ps, average=mode)()
    assert all(torch.allclose(p.grad, torch.full_like(p, expect)) for p in ps)
torch.cuda.synchronize()
dist.barrier()
print("rank", rank, "ok")
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_paths(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, CP_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29621", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
