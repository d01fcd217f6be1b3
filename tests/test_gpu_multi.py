"""Multi-rank parity: one process per rank under torchrun.  With >= 2 GPUs on the box the ranks sit on separate
GPUs and talk NCCL; on a 1-GPU box (the driver's test box) the SAME workers run as two ranks on cuda:0 over
gloo (CP_DIST_BACKEND=gloo; dist.init_from_env folds LOCAL_RANK onto the GPUs that exist), so the sharded
code paths -- SyncBN, the sharded batch x batch head, the flat gradient all-reduce, sharded evaluation,
train.main under WORLD_SIZE 2, the per-rank fold split -- are exercised on the final code either way.
The sharded computation must reproduce the single-GPU one on the concatenated batch."""
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
nccl = dist.get_backend() == "nccl"

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

def mark(what):
    torch.cuda.synchronize()
    sys.stderr.write(f"[rank {rank}] {what}\n"); sys.stderr.flush()

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

mark("clip head ok")
# ---- flat-bucket gradient all-reduce over NCCL: average and sum
ps = [torch.nn.Parameter(torch.zeros(1000, 3, device=dev)), torch.nn.Parameter(torch.zeros(7, device=dev))]
for mode, expect in ((True, (world + 1) / 2.0), (False, world * (world + 1) / 2.0)):
    for p in ps:
        p.grad = torch.full_like(p, float(rank + 1))
    cpdist.FlatGradAllReduce(ps, average=mode)()
    assert all(torch.allclose(p.grad, torch.full_like(p, expect)) for p in ps)
mark("flat all-reduce ok")
# ---- SyncBN: world ranks x (B/world) groups with statistics over every rank's rows == one rank x B groups
from contrastiveprosthetics_b200.models import Model
params = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 0.0, 'reg_glove': 0.0}
Bg = 24 * world
g = torch.Generator().manual_seed(5)
EMG_all = (torch.randn(Bg, 41, 1, 1, 12, generator=g) + 0.5 * torch.randn(1, 41, 1, 1, 12, generator=g)).to(dev)
label = torch.arange(41, device=dev).repeat(Bg)
for adabn in (True, False):
    torch.manual_seed(42)
    m_sh = Model(dict(params), adabn=adabn, device=str(dev))
    torch.manual_seed(42)
    m_full = Model(dict(params), adabn=adabn, device=str(dev))
    m_sh.set_train(); m_full.set_train()
    m_sh.emg_net.sync_bn = True
    per = Bg // world
    lo = m_sh.loss(m_sh.forward(EMG_all[rank * per:(rank + 1) * per], None, label[:per * 41]), label[:per * 41])
    lo.backward()
    cpdist.FlatGradAllReduce(list(m_sh.parameters()))()
    lsum = lo.detach().clone(); dist.all_reduce(lsum); lsum /= world
    lf = m_full.loss(m_full.forward(EMG_all, None, label), label)
    lf.backward()
    assert abs(lsum.item() - lf.item()) < 2e-6 * abs(lf.item()), (lsum.item(), lf.item())
    worst = 0.0
    for (k, a), (_, b) in zip(m_sh.named_parameters(), m_full.named_parameters()):
        if b.grad is None:
            continue
        worst = max(worst, rel(a.grad, b.grad))
    # float partial sums are tiled differently in the two runs, so scale/shift can differ by an ulp and a
    # pre-activation within rounding noise of 0 may take the other ReLU branch: one flip moves a gradient
    # tensor by ~1/sqrt(#elements) of its norm (DESIGN.md "ReLU kinks"); without flips the match is ~1e-6
    assert worst < 2e-2, worst
    if not adabn:       # running statistics follow the GLOBAL batch
        for (k, a), (_, b) in zip(m_sh.named_buffers(), m_full.named_buffers()):
            if a.dtype.is_floating_point:
                assert rel(a, b) < 1e-6, k
    # local-BN (default) differs from the global-batch result: the switch really changes the semantics
    torch.manual_seed(42)
    m_loc = Model(dict(params), adabn=adabn, device=str(dev)); m_loc.set_train()
    ll = m_loc.loss(m_loc.forward(EMG_all[rank * per:(rank + 1) * per], None, label[:per * 41]), label[:per * 41])
    lsum2 = ll.detach().clone(); dist.all_reduce(lsum2); lsum2 /= world
    assert abs(lsum2.item() - lf.item()) > 1e-5 * abs(lf.item())
    print("rank", rank, "adabn", adabn, "syncbn worst grad rel", worst)
mark("syncbn ok")
# ---- step.LeanTrainStep with the gradient bucket averaged in place == the autograd step + FlatGradAllReduce + Adam;
#      (NCCL) its CUDA-graph capture, all-reduce inside the graph, == the eager lean step
from contrastiveprosthetics_b200.step import LeanTrainStep
from contrastiveprosthetics_b200.graph import GraphedTrainStep
hp = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-4, 'reg_glove': 1e-3}
per = 8
g = torch.Generator().manual_seed(9)
steps = [(torch.randn(per * world, 41, 1, 1, 12, generator=g) + 0.5 * torch.randn(1, 41, 1, 1, 12, generator=g)).to(dev)
         for _ in range(4)]
mine = [b[rank * per:(rank + 1) * per].contiguous() for b in steps]
lab = torch.arange(41, device=dev).repeat(per)
def fresh():
    torch.manual_seed(42)
    m = Model(dict(hp), adabn=True, device=str(dev)); m.set_train()
    return m
m_ref = fresh()
opts = [torch.optim.Adam(m_ref.emg_net.parameters(), lr=1e-3), torch.optim.Adam(m_ref.glove_net.parameters(), lr=3e-3)]
sync = cpdist.FlatGradAllReduce(list(m_ref.emg_net.parameters()) + list(m_ref.glove_net.parameters()))
for EMG in mine:
    total = m_ref.loss(m_ref.forward(EMG, None, lab), lab) + m_ref.l2()
    for o in opts: o.zero_grad(set_to_none=True)
    total.backward(); sync()
    for o in opts: o.step()
mark("autograd reference steps ok")
m_lean = fresh()
lean = LeanTrainStep(m_lean, 1e-3, 3e-3, sync_grads=True)
for EMG in mine:
    lean(EMG)
# Adam turns rounding-level differences of single gradients (here: avg(g + l2) vs avg(g) + l2) into lr-sized differences
# of single parameters (tests/test_gpu_step.py::test_lean_step_equals_autograd_step): compare in the mean, and the loss
worst, n_par = 0.0, 0
for (k, a), (_, b) in zip(m_lean.state_dict().items(), m_ref.state_dict().items()):
    if a.dtype.is_floating_point:
        worst += float((a - b).abs().sum()); n_par += a.numel()
        other = a.clone(); dist.broadcast(other, src=0)
        assert torch.equal(a, other), ("replicas diverged", k)
worst /= n_par
assert worst < 1e-5, worst
with torch.no_grad():
    l_a = m_lean.loss(m_lean.forward(mine[0], None, lab), lab).item()
    l_b = m_ref.loss(m_ref.forward(mine[0], None, lab), lab).item()
assert abs(l_a - l_b) < 1e-3 * abs(l_b), (l_a, l_b)
mark("eager lean steps ok")
if nccl:
    m_g = fresh()
    o_g = [torch.optim.Adam(m_g.emg_net.parameters(), lr=1e-3), torch.optim.Adam(m_g.glove_net.parameters(), lr=3e-3)]
    gs = GraphedTrainStep(m_g, o_g, mine[0], sync_grads=True, lean=True)
    for EMG in mine:
        gs(EMG)
    for (k, a), (_, b) in zip(m_g.state_dict().items(), m_lean.state_dict().items()):
        assert torch.equal(a, b), ("graphed lean step != eager lean step", k)
    # a CUDA graph that holds captured NCCL collectives must go before the communicator does (destroy_process_group
    # otherwise waits forever -- seen on 2 x B200)
    del gs
    import gc; gc.collect()
print("rank", rank, "lean step vs autograd + all-reduce: mean |dp|", worst, "loss", l_a, l_b)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.write(f"rank {rank} ok\n"); sys.stdout.flush()          # one write: the ranks share the pipe
import threading
threading.Timer(30.0, lambda: os._exit(0)).start()                  # every check has passed: never hang in teardown
try:
    dist.destroy_process_group()
except Exception as e:                                              # (a peer that is already gone)
    sys.stderr.write(f"[rank {rank}] destroy_process_group: {e!r}\n")
sys.stderr.flush()
os._exit(0)
'''


_TRAIN_WORKER = r'''
import os, sys, numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CP_ROOT"])
from contrastiveprosthetics_b200 import dist as cpdist, train as T
tmp = os.environ["CP_TMP"]
argv = ["--final_epochs=1", "--crossval_size=3", "--crossval_epochs=1", "--batch_size=24", "--test", "--synthetic",
        "--no_verbose", f"--data_dir={tmp}/data/", f"--checkpoint_dir={tmp}/ckpt/"]
args = T.build_parser().parse_args(argv)
# keep the model train.main builds for the checks below
kept = {}
orig = T.train_loop
def spy(*a, **k):
    out = orig(*a, **k)
    kept["model"] = out[1]
    return out
T.train_loop = spy
loss, acc = T.main(args)
rank, world = cpdist.rank(), cpdist.world_size()
assert world == 2
dev = torch.device("cuda", torch.cuda.current_device())
# (1) the replicas are ONE model after training: identical parameters and buffers on every rank
m = kept["model"]
for name, t in list(m.named_parameters()) + list(m.named_buffers()):
    ref = t.detach().clone()
    cpdist.broadcast_(ref)
    assert torch.equal(ref, t.detach()), name
# (2) sharded evaluation reports the same numbers on every rank ...
r = torch.tensor([loss, acc], dtype=torch.float64, device=dev)
r0 = r.clone(); cpdist.broadcast_(r0)
assert torch.equal(r, r0), (r, r0)
# (3) ... and they are the numbers of an UN-sharded evaluation of the same model on the same batches
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.utils import TaskWrapper
ds = DB23(db2=False, device=dev); ds.load_synthetic()
tw = TaskWrapper(ds)
T.shuff = False
torch.manual_seed(7); torch.cuda.manual_seed(7)          # same per-class window draws in both evaluations
sh_loss, sh_acc = T.test(m, tw)
sh_vote, sh_pred = m.voting_raw().copy(), m.y_pred_raw().copy()
saved = (cpdist.rank, cpdist.world_size)
cpdist.rank, cpdist.world_size = (lambda: 0), (lambda: 1)
try:
    torch.manual_seed(7); torch.cuda.manual_seed(7)
    one_loss, one_acc = T.test(m, tw)
    one_vote, one_pred = m.voting_raw().copy(), m.y_pred_raw().copy()
finally:
    cpdist.rank, cpdist.world_size = saved
assert abs(sh_loss - one_loss) < 1e-5 * abs(one_loss), (sh_loss, one_loss)
assert sh_vote.shape == one_vote.shape and sh_pred.shape == one_pred.shape
# integer decisions: identical unless a logit pair sits within rounding noise (statistics are reduced in another order)
assert (sh_pred != one_pred).mean() < 1e-3 and abs(sh_acc - one_acc) < 1e-3, (sh_acc, one_acc)
# (4) cross-validation folds were split per rank and gathered: both ranks hold the full table, and each fold's row
#     is the one its owner computed
vals = np.load(f"{tmp}/data/cross_val_values.npy"); keys = np.load(f"{tmp}/data/cross_val_keys.npy")
assert vals.shape == (3, 2) and keys.shape == (3, 7) and np.isfinite(vals).all()
sys.stdout.write(f"rank {rank} ok {loss} {acc}\n"); sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
'''


def _run_two_ranks(tmp_path, worker, port, timeout=420):
    script = tmp_path / "w.py"
    script.write_text(worker)
    env = dict(os.environ, CP_ROOT=ROOT, CP_TMP=str(tmp_path))
    if torch.cuda.device_count() < 2:
        env["CP_DIST_BACKEND"] = "gloo"             # two ranks on the one GPU
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" ok") == 2, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_two_rank_sharded_paths(tmp_path):
    """SyncBN == global batch, sharded batch x batch head == single GPU, flat gradient all-reduce."""
    _run_two_ranks(tmp_path, _WORKER, 29621)


def test_two_rank_train_main(tmp_path):
    """train.main under WORLD_SIZE 2: identical replicas after training, sharded evaluation == un-sharded
    evaluation, folds split per rank."""
    _run_two_ranks(tmp_path, _TRAIN_WORKER, 29623)
