"""Host-side logic that runs without a GPU: split masks / D / len of DB23, trial generation and
sharding, the C-ABI surface, loud failure without CUDA, world_size-2 gloo paths."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import _lib, subset
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.synthetic import synth_emg
from contrastiveprosthetics_b200.utils import TaskWrapper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gd(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset.npz"))


@pytest.fixture(scope="module")
def emg():
    return synth_emg()


@pytest.mark.parametrize("db2", [False, True])
def test_db23_splits_match_reference(gd, emg, db2):
    ds = DB23(db2=db2, device="cpu")
    ds.load_tensors(emg)
    for split in ("train", "val", "test"):
        getattr(ds, "set_" + split)()
        tag = f"db2{int(db2)}_{split}"
        assert ds.D == int(gd[tag + "_D"])
        assert len(ds) == int(gd[tag + "_len"])
        assert len(TaskWrapper(ds)) == int(gd[tag + "_twlen"])
        assert np.array_equal(ds.tasks_mask.numpy(), gd[tag + "_tasks"])
        assert np.array_equal(ds.people_mask.numpy(), gd[tag + "_people"])
        assert np.array_equal(ds.rep_mask.numpy(), gd[tag + "_reps"])
        assert np.array_equal(ds.EMG_use[gd[tag + "_rows"]].numpy(), gd[tag + "_EMG_use"])
        assert np.array_equal(ds.tensor[gd[tag + "_trows"]].numpy(), gd[tag + "_tensor"])


def test_db23_mixed_subjects_split(emg):
    """Config 3: DB2 + DB3 subjects mixed (46 people, DB3 repetition split); the 6 DB3 subjects are
    11-channel (channel 10 zeroed).  Checked against the oracle restatement."""
    from oracle import dataset as OD
    ds = DB23(mixed=True, device="cpu")
    ds.load_tensors(emg)
    E = emg.transpose(0, 1).numpy()
    for split, D in (("train", 46 * 3 * 100), ("val", 46 * 1 * 4), ("test", 46 * 2 * 4)):
        getattr(ds, "set_" + split)()
        assert ds.D == D and ds.PEOPLE == 46 and len(TaskWrapper(ds)) == D
        use, tensor, D0 = OD.load_valid(E, False, split, mixed=True)
        assert D0 == D
        assert np.array_equal(ds.EMG_use.numpy(), use)
        assert np.array_equal(ds.tensor.numpy(), tensor)
    ds.set_train()
    rows = ds.EMG_use.reshape(41, 46, 3, 100, 12)
    assert float(rows[:, 40:, :, :, 10].abs().max()) == 0.0 and float(rows[:, :40, :, :, 10].abs().min()) > 0.0
    assert float(rows[:, 40:, :, :, 11].abs().min()) > 0.0


def test_no_cpu_fallback(emg):
    """Indexing is the CUDA gather; on CPU tensors it must raise, not silently index with torch."""
    ds = DB23(db2=False, device="cpu")
    ds.load_tensors(emg)
    ds.set_train()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ds[torch.arange(41)]
    from contrastiveprosthetics_b200.models import Model
    m = Model({'d_e': 16, 'dp_emg': 0., 'dp_glove': 0., 'reg_emg': 0., 'reg_glove': 0.}, device="cpu")
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(2, 41, 1, 1, 12), torch.zeros(2, 41, 20), torch.arange(41).repeat(2))


@pytest.mark.parametrize("adabn,n_reg", [(True, 12), (False, 21)])
def test_lean_step_host_layout(adabn, n_reg):
    """step.LeanTrainStep's host side (no compute call without a GPU): the regularised set is the reference's name filter
    (models.py:344-349 / 467-472: 10 + 2 tensors with AdaBN, 19 + 2 with --no_adabn), every parameter's .grad
    is a 512-byte-aligned view of ONE flat bucket in optimizer order, lr / reg follow the two Adams of train.py:72-73,
    and the step itself refuses CPU tensors."""
    from contrastiveprosthetics_b200.models import Model
    from contrastiveprosthetics_b200.step import LeanTrainStep, from_optimizers
    from oracle import model as OM
    hp = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0., 'reg_emg': 1e-4, 'reg_glove': 1e-3}
    torch.manual_seed(42)
    m = Model(dict(hp), adabn=adabn, device="cpu")
    s = LeanTrainStep(m, 1e-3, 3e-3)
    named = [("emg_net." + k, p) for k, p in m.emg_net.named_parameters()] + \
            [("glove_net." + k, p) for k, p in m.glove_net.named_parameters()]
    assert [id(p) for _, p in named] == [id(p) for p in s.params]
    assert sum(p.numel() for p in s.params) == 2_027_616          # SURVEY a12 (logit_scale untouched)
    assert s._n_reg == n_reg
    reg_names = {k for (k, _), ni in zip(named, s._norm_index) if ni >= 0}
    assert reg_names == {k for k, _ in named if "bn" not in k and "bias" not in k}
    # the oracle's penalty walks the same keys: its value changes iff a regularised tensor changes
    sd = {k: p.detach().clone() for k, p in named}
    base = float(OM.l2_penalty(sd, 1.0, 1.0))
    for k in sd:
        sd2 = dict(sd)
        sd2[k] = sd[k] * 2 + 1
        assert (abs(float(OM.l2_penalty(sd2, 1.0, 1.0)) - base) > 0) == (k in reg_names), k
    off = 0
    for (k, p), g, o, li, rg in zip(named, s.grad_views, s._offs, s._lr_index, s._reg):
        assert o == off and o % 128 == 0 and p.grad is g and g.shape == p.shape
        assert g.data_ptr() == s.grad_flat.data_ptr() + 4 * o
        assert li == (0 if k.startswith("emg_net.") else 1)
        assert rg == pytest.approx(1e-4 if k.startswith("emg_net.") else 1e-3)
        off += (p.numel() + 127) // 128 * 128
    assert off == s.numel == s.exp_avg.numel() == s.exp_avg_sq.numel()
    assert m.emg_net.dropout_step.data_ptr() == s.counters.data_ptr()
    m.set_train()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        s(torch.zeros(2, 41, 1, 1, 12))
    with pytest.raises(RuntimeError, match="expected a"):
        s(torch.zeros(2, 40, 1, 1, 12))
    # hyper-parameters are read off the two torch optimizers; anything but train.py:72-73's plain Adam is refused
    o_e, o_g = torch.optim.Adam(m.emg_net.parameters(), lr=2e-3), torch.optim.Adam(m.glove_net.parameters(), lr=5e-3)
    s2 = from_optimizers(m, [o_e, o_g])
    assert s2.lr.tolist() == [2e-3, 5e-3] and s2.betas == (0.9, 0.999) and s2.eps == 1e-8
    with pytest.raises(RuntimeError, match="weight_decay"):
        from_optimizers(m, [torch.optim.Adam(m.emg_net.parameters(), lr=1e-3, weight_decay=1e-2), o_g])
    with pytest.raises(NotImplementedError):
        LeanTrainStep(Model(dict(hp), prediction=True, device="cpu"), 1e-3, 1e-3)


def test_cabi_exports_every_declared_symbol():
    """libcpros.so loads and exports exactly what include/cpros.h declares."""
    _lib.build()
    hdr = open(os.path.join(ROOT, "include", "cpros.h")).read()
    declared = set(re.findall(r"\b(cp_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cp_encoder_tensors", "cp_encoder_opts", "cp_glove_tensors", "cp_glove_opts", "cp_allreduce_fn"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().cp_version() >= 100
    # struct mirrors: 4 + 7 + 7 + 1 + 4*9 pointers
    assert ctypes.sizeof(_lib.EncoderTensors) == 8 * (4 + 14 + 1 + 36)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of every struct in include/cpros.h, as gcc lays them out, == the ctypes mirrors."""
    fields = {"cp_encoder_tensors": ["conv1_w", "fc_w", "proj_w", "bn_w", "bn_rv"],
              "cp_encoder_opts": ["bn_mode", "engine", "bn_momentum", "bn_eps", "dropout_p", "save_for_backward",
                                  "dropout_seed", "ext_masks", "dropout_step", "allreduce", "allreduce_user", "trunk_only"],
              "cp_cls_tensors": ["w1", "b1", "bn_w", "bn_b", "bn_rm", "bn_rv", "w2"],
              "cp_glove_tensors": ["w0", "bn0_b", "w", "b", "bn_w", "bn_b", "proj_w"],
              "cp_glove_opts": ["glove_dim", "save_for_backward", "bn_eps", "dropout_p", "dropout_seed", "ext_masks",
                                "dropout_step"]}
    mirrors = {"cp_encoder_tensors": _lib.EncoderTensors, "cp_encoder_opts": _lib.EncoderOpts,
               "cp_glove_tensors": _lib.GloveTensors, "cp_glove_opts": _lib.GloveOpts, "cp_cls_tensors": _lib.ClsTensors}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cpros.h"', 'int main(void) {']
    for st, fs in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for st, fs in fields.items():
        assert int(got[st]) == ctypes.sizeof(mirrors[st]), st
        for f in fs:
            assert int(got[f"{st}.{f}"]) == getattr(mirrors[st], f).offset, (st, f)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_make_trials_and_sharding():
    masks, sizes = subset.make_trials(sizes=range(1, 41), trials_per_size=144, seed=0)
    assert masks.shape == (40 * 144, 41) and masks[:, 40].all()
    assert np.array_equal(masks.sum(1), sizes + 1)
    # size 40 == the full 41-class set (data/min_grasp.xlsx: min == max at size 40)
    assert masks[sizes == 40].all()
    cover = []
    for r in range(8):
        lo, hi = subset.shard_trials(144, r, 8)
        assert hi - lo == 18
        cover += list(range(lo, hi))
    assert cover == list(range(144))
    lo, hi = subset.shard_trials(5, 7, 8)
    assert lo == hi


def test_batches_cover_every_item_once_per_rank_split(emg):
    ds = DB23(db2=False, device="cpu")
    ds.load_tensors(emg)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_val()
    seen = []
    tw.get_batch = lambda items: seen.append(items.clone())          # host logic only
    for w in range(2):
        g = torch.Generator().manual_seed(3)
        list(tw.batches(5, shuffle=True, generator=g, rank=w, world_size=2))
    allitems = torch.cat(seen).sort().values
    assert torch.equal(allitems, torch.arange(ds.D))


@pytest.mark.parametrize("D,batch,world", [(160, 64, 8), (9, 9, 8), (100, 33, 8), (48, 8, 3), (17, 5, 4), (7, 64, 8),
                                            (13800, 4096, 8), (24, 7, 2)])
def test_every_rank_runs_the_same_number_of_batches(D, batch, world):
    """Sample sharding must not let a rank skip a step the others run (its all-reduces would never be matched):
    every global batch is dealt out evenly, and a ragged tail smaller than the world is dropped on ALL ranks."""
    class _DS:
        device = torch.device("cpu")
    tw = TaskWrapper.__new__(TaskWrapper)
    tw.__dict__["dataset"] = _DS()
    tw.__dict__["device"] = torch.device("cpu")
    TaskWrapper.__len__ = TaskWrapper.__len__          # (len() goes through the class)
    tw.__dict__["dataset"].D = D
    plans = []
    for r in range(world):
        g = torch.Generator().manual_seed(11)
        plans.append(tw.batch_plan(batch, shuffle=True, generator=g, rank=r, world_size=world))
    counts = {len(p) for p in plans}
    assert len(counts) == 1, counts
    seen = []
    for b in range(len(plans[0])):
        n_global = plans[0][b][2]
        assert n_global >= world
        sizes = [p[b][0].numel() for p in plans]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1 and sum(sizes) == n_global
        los = [p[b][1] for p in plans]
        assert los == [sum(sizes[:r]) for r in range(world)]
        seen.append(torch.cat([p[b][0] for p in plans]))
    covered = torch.cat(seen).sort().values if seen else torch.zeros(0, dtype=torch.long)
    tail = D % batch
    dropped = tail if 0 < tail < world else 0
    assert covered.numel() == D - dropped and covered.unique().numel() == covered.numel()


_GLOO_WORKER = r'''
import os, sys, torch, numpy as np
sys.path.insert(0, os.environ["CP_ROOT"])
from contrastiveprosthetics_b200 import dist as cpdist, subset
rank, world, dev = cpdist.init_from_env("gloo")
assert world == 2 and dev.type == "cpu"
# replicas start from ONE model: rank 0's parameters and buffers everywhere (train.train_loop)
torch.manual_seed(100 + rank)
net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
net[1].running_mean.fill_(float(rank)); net[1].num_batches_tracked.fill_(7 * rank + 1)
cpdist.broadcast_module(net)
ref = [t.clone() for t in list(net.parameters()) + list(net.buffers())]
for t in ref:
    cpdist.broadcast_(t)                     # rank 0's copy of what this rank now holds
assert all(torch.equal(a, b) for a, b in zip(ref, list(net.parameters()) + list(net.buffers())))
assert float(net[1].running_mean[0]) == 0.0 and int(net[1].num_batches_tracked) == 1
# sharded evaluation: per-group integer arrays of a global batch put back together (exact), uneven shards
n_global = 7
lo, hi = cpdist.even_shard(n_global)
full = torch.arange(n_global * 3, dtype=torch.int32).reshape(n_global, 3) + 1
got = cpdist.assemble_rows(full[lo:hi].clone(), lo, n_global)
assert torch.equal(got, full)
# the row collectives the batch x batch head needs, on a backend without tensor collectives
x = torch.full((2, 3), float(rank + 1))
assert torch.equal(cpdist.all_gather_rows(x), torch.tensor([[1.0] * 3] * 2 + [[2.0] * 3] * 2))
part = torch.arange(12, dtype=torch.float32).reshape(4, 3) * (rank + 1)
assert torch.equal(cpdist.reduce_scatter_rows(part.clone()), 3 * torch.arange(12, dtype=torch.float32).reshape(4, 3)[rank * 2:rank * 2 + 2])
# one permutation per job whatever the ranks' generators say
from contrastiveprosthetics_b200.utils import TaskWrapper
class _DS:
    device = torch.device("cpu"); D = 23
tw = TaskWrapper.__new__(TaskWrapper); tw.__dict__["dataset"] = _DS(); tw.__dict__["device"] = torch.device("cpu")
plan = tw.batch_plan(6, shuffle=True, generator=torch.Generator().manual_seed(50 + rank), rank=rank, world_size=world)
mine = torch.cat([p[0] for p in plan])
both = cpdist.assemble_rows(mine.clone(), 0 if rank == 0 else 23 - mine.numel(), 23)
assert both.sort().values.tolist() == list(range(23)), both
# flat-bucket gradient averaging
torch.manual_seed(0)
ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
for i, p in enumerate(ps):
    p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
cpdist.FlatGradAllReduce(ps)()
assert torch.allclose(ps[0].grad, torch.full((5, 3), 1.5)) and torch.allclose(ps[1].grad, torch.full((7,), 3.0))
# the lean step's gradient bucket is averaged IN PLACE (one all-reduce, no pack / unpack): every parameter's .grad is a
# view of it, so each rank then sees the mean of the ranks' gradients (step.LeanTrainStep._all_reduce)
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.step import LeanTrainStep
torch.manual_seed(42)
lean = LeanTrainStep(Model({'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0., 'reg_emg': 1e-5, 'reg_glove': 1e-5}, device="cpu"),
                     1e-3, 1e-3, sync_grads=True)
assert lean.sync_grads
for i, g in enumerate(lean.grad_views):
    g.fill_(float((rank + 1) * (i + 1)))
lean._all_reduce()
for i, p in enumerate(lean.params):
    assert torch.equal(p.grad, torch.full_like(p, 1.5 * (i + 1))), i
# trial sharding + exact integer reduction
masks, _ = subset.make_trials(sizes=[3], trials_per_size=9, seed=1)
lo, hi = subset.shard_trials(len(masks), rank, world)
correct = torch.zeros(len(masks), dtype=torch.int64); correct[lo:hi] = torch.arange(lo, hi) + 1
cpdist.sum_counts(correct)
assert torch.equal(correct, torch.arange(len(masks)) + 1)
print("rank", rank, "ok")
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, CP_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


# ---- batch x batch (CLIP) head: the N > 1 exchange logic of clip.clip_head under gloo.  The six libcpros entry
# points are replaced by a torch-CPU test double (tests only -- the product has no such path), so what is tested
# is the partitioning: all-gather of Ghat, all-reduce of the column sums, reduce-scatter of d Ghat, global
# loss / count, gradients of the GLOBAL loss w.r.t. the rank's own embeddings.
_GLOO_CLIP_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["CP_ROOT"])
from contrastiveprosthetics_b200 import clip as C, dist as cpdist
from oracle import clip as OC

class TorchOps:
    @staticmethod
    def normalize(x):
        inv = 1.0 / x.norm(dim=1)
        return x * inv[:, None], inv
    @staticmethod
    def transpose(xhat):
        n = xhat.shape[0]; ld = (n + 3) // 4 * 4
        out = torch.zeros(16, ld, dtype=xhat.dtype); out[:, :n] = xhat.t(); return out
    @staticmethod
    def _e(own, loop_t, n_loop, scale):
        return torch.exp(scale * (own @ loop_t[:, :n_loop] - 1.0))
    @staticmethod
    def sums(own, loop_t, n_loop, scale, want_argmax):
        S = own @ loop_t[:, :n_loop]
        return TorchOps._e(own, loop_t, n_loop, scale).sum(1), (S.argmax(1).to(torch.int32) if want_argmax else None)
    @staticmethod
    def loss(ehat, ghat, rowsum, colsum, B, scale, row_arg, row0):
        n = ehat.shape[0]
        t = (rowsum.log() + colsum.log() + 2 * scale - 2 * scale * (ehat * ghat).sum(1)).sum() / (2 * B)
        return t, (row_arg.long() == torch.arange(row0, row0 + n)).sum().to(torch.int32)
    @staticmethod
    def grad(own, loop_t, n_loop, scale, own_sum, loop_sum, coef):
        c = TorchOps._e(own, loop_t, n_loop, scale) * (1.0 / own_sum[:, None] + 1.0 / loop_sum[None, :])
        return coef * (c @ loop_t[:, :n_loop].t())
    @staticmethod
    def embed_backward(d_hat, xhat, other_hat, inv_norm, diag_coef):
        dh = d_hat - diag_coef * other_hat
        return (dh - xhat * (xhat * dh).sum(1, keepdim=True)) * inv_norm[:, None]

for name in ("normalize", "transpose", "sums", "loss", "grad", "embed_backward"):      # test-side patch of the six entry points
    setattr(C._CudaOps, name, staticmethod(getattr(TorchOps, name)))
rank, world, dev = cpdist.init_from_env("gloo")
n = 5
g = torch.Generator().manual_seed(0)
E_all = torch.randn(world * n, 16, generator=g, dtype=torch.float64)
G_all = torch.randn(world * n, 16, generator=g, dtype=torch.float64) + 0.5 * E_all
for logit_scale in (0.0, 1.2):
    E = E_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    G = G_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    loss, ncor, arg = C.clip_head(E, G, logit_scale)
    loss.backward()
    ref_loss, dE, dG = OC.clip_loss_sharded(list(E_all.split(n)), list(G_all.split(n)), logit_scale)
    full = OC.clip_loss(E_all, G_all, logit_scale)
    assert abs(loss.item() - ref_loss.item()) < 1e-12, (loss.item(), ref_loss.item())
    assert torch.allclose(E.grad, dE[rank], rtol=1e-10, atol=1e-14)
    assert torch.allclose(G.grad, dG[rank], rtol=1e-10, atol=1e-14)
    assert int(ncor) == full["n_correct"]
    assert torch.equal(arg.long(), full["pred"][rank * n:(rank + 1) * n])
# gradient SUM (not average) for a loss normalised by the global batch
p = torch.nn.Parameter(torch.zeros(3)); p.grad = torch.full((3,), float(rank + 1))
cpdist.FlatGradAllReduce([p], average=False)()
assert torch.allclose(p.grad, torch.full((3,), 3.0))
print("rank", rank, "ok")
'''


def test_gloo_clip_head_exchange(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_CLIP_WORKER)
    env = dict(os.environ, CP_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29612", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_subjects_of_rows(emg):
    """DB23.subjects_of (per-subject AdaBN, models.py:245): row id = class*D + k, k over (person, repetition, window) of
    the split (load.py:233-251) -> the person's index on the 46-subject axis, in train and in the voted evaluation."""
    ds = DB23(device="cpu", mixed=True)
    ds.load_tensors(emg)
    for setter, per_rep in ((ds.set_train, 100), (ds.set_test, 4)):
        setter()
        D, P, R = ds.D, ds.PEOPLE, ds.REPS
        assert D == P * R * per_rep
        rows = torch.tensor([0, per_rep * R - 1, per_rep * R, D - 1, D, 5 * D + 3 * per_rep * R + 7, 41 * D - 1])
        k = rows % D
        expect = ds.people_mask[k // (per_rep * R)]
        assert torch.equal(ds.subjects_of(rows), expect)
        # and against the data itself: the gathered row IS that person's window
        sub = ds.EMG[ds.tasks_mask][:, ds.people_mask][:, :, ds.rep_mask][:, :, :, :100]
        r = int(rows[5])
        c, kk = r // D, r % D
        p, rem = kk // (per_rep * R), kk % (per_rep * R)
        assert int(ds.subjects_of(torch.tensor([r]))[0]) == int(ds.people_mask[p])
        assert rem // per_rep < R and c == 5
