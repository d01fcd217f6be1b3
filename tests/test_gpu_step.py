"""step.LeanTrainStep (train.py:95-108 without autograd / torch.optim): cp_step_prologue and cp_adam_step against
cp_l2_forward and torch.optim.Adam, the whole step against the autograd step, its CUDA-graph capture against itself."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-4, 'reg_glove': 1e-3, 'lr_emg': 1e-3, 'lr_glove': 3e-3}


def _model(dp=0.0, adabn=True):
    from contrastiveprosthetics_b200.models import Model
    torch.manual_seed(42)
    m = Model(dict(PARAMS, dp_emg=dp), adabn=adabn, device="cuda")
    m.set_train()
    return m


def _batches(n, B=8):
    g = torch.Generator().manual_seed(1)
    return [(torch.randn(B, 41, 1, 1, 12, generator=g) + 0.5 * torch.randn(1, 41, 1, 1, 12, generator=g)).cuda()
            for _ in range(n)]


def test_prologue_norms_equal_l2_forward_and_counters_advance():
    from contrastiveprosthetics_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    ts = [torch.randn(s, device="cuda", generator=g) * sc for s, sc in
          [(576, 1.0), (36864, 0.05), (512 * 768, 0.03), (512 * 512, 1e-3), (16 * 512, 2.0), (7, 1.0), (16 * 41, 0.3)]]
    ts.append(torch.zeros(100, device="cuda"))                       # a zero norm
    n = len(ts)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
    sizes = (ctypes.c_int64 * n)(*[t.numel() for t in ts])
    ref = torch.empty(n, device="cuda")
    tot = torch.empty((), device="cuda")
    nb = L.cp_l2_workspace_bytes(n)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    _lib.check(L.cp_l2_forward(ptrs, sizes, n, _lib.ptr(ref), _lib.ptr(tot), _lib.ptr(ws), nb, _lib.stream()))
    norms = torch.full((n,), -1.0, device="cuda")
    counters = torch.tensor([5, 0], dtype=torch.int64, device="cuda")
    nb2 = L.cp_step_workspace_bytes(n)
    ws2 = torch.zeros(nb2, dtype=torch.uint8, device="cuda")
    for rep in range(3):                                             # the ticket resets itself
        norms.fill_(-1.0)
        _lib.check(L.cp_step_prologue(ptrs, sizes, n, _lib.ptr(norms), _lib.ptr(counters), 2, _lib.ptr(ws2), nb2,
                                      _lib.stream()))
        assert torch.equal(norms, ref), rep
        assert counters.tolist() == [6 + rep, 1 + rep]
    assert torch.allclose(ref[:-1], torch.stack([t.double().norm().float() for t in ts[:-1]]), rtol=1e-6)
    assert ref[-1].item() == 0.0


@pytest.mark.parametrize("fused", [False, True])
def test_adam_step_matches_torch_adam(fused):
    """Flat-bucket Adam (+ the regulariser's gradient) against torch.optim.Adam on explicit gradients."""
    from contrastiveprosthetics_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 9), (64,), (512, 768), (512,), (16, 512), (1031,), (3,)]
    lrs = [1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 3e-2, 3e-2]
    lr_index = [0, 0, 0, 0, 0, 1, 1]
    regs = [1e-2, 0.0, 1e-2, 0.0, 1e-2, 0.5, 0.0]
    norm_index = [0, -1, 1, -1, 2, 3, -1]
    ps = [torch.randn(s, device="cuda", generator=g) * 0.1 for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opts = [torch.optim.Adam([r for r, i in zip(ref, lr_index) if i == k], lr=lr, fused=fused) for k, lr in ((0, 1e-3), (1, 3e-2))]
    n = len(ps)
    offs, off = [], 0
    for p in ps:
        offs.append(off)
        off += (p.numel() + 127) // 128 * 128
    grads = torch.zeros(off, device="cuda")
    m = torch.zeros(off, device="cuda")
    v = torch.zeros(off, device="cuda")
    lr = torch.tensor([1e-3, 3e-2], dtype=torch.float64, device="cuda")
    reg_list = [i for i in range(n) if norm_index[i] >= 0]
    norms = torch.zeros(len(reg_list), device="cuda")
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")
    ws = torch.zeros(L.cp_step_workspace_bytes(len(reg_list)), dtype=torch.uint8, device="cuda")
    A = lambda ct, xs: (ct * len(xs))(*xs)                                                # noqa: E731
    p_ptrs, sizes, offs_c = A(ctypes.c_void_p, [p.data_ptr() for p in ps]), A(ctypes.c_int64, [p.numel() for p in ps]), A(ctypes.c_int64, offs)
    r_ptrs, r_sizes = A(ctypes.c_void_p, [ps[i].data_ptr() for i in reg_list]), A(ctypes.c_int64, [ps[i].numel() for i in reg_list])
    worst = 0.0
    for step in range(8):
        gs = [torch.randn(s, device="cuda", generator=g) * (10.0 ** -(step % 4)) for s in shapes]
        gs[3][::7] = 0.0                                                                 # exact zeros stay put
        for o, gg, p in zip(offs, gs, ps):
            grads[o:o + p.numel()] = gg.reshape(-1)
        for r, gg, rg, ni in zip(ref, gs, regs, norm_index):
            nt = r.detach().double().norm().float()
            r.grad = gg + (torch.tensor(rg, device="cuda") / nt) * r.detach() if ni >= 0 else gg.clone()
        for o in opts:
            o.step()
        _lib.check(L.cp_step_prologue(r_ptrs, r_sizes, len(reg_list), _lib.ptr(norms), _lib.ptr(counters), 2, _lib.ptr(ws),
                                      ws.numel(), _lib.stream()))
        _lib.check(L.cp_adam_step(p_ptrs, sizes, offs_c, n, _lib.ptr(grads), _lib.ptr(m), _lib.ptr(v), _lib.ptr(lr),
                                  A(ctypes.c_int32, lr_index), A(ctypes.c_float, regs), A(ctypes.c_int32, norm_index),
                                  _lib.ptr(norms), _lib.ptr(counters[1:2]), 0.9, 0.999, 1e-8, _lib.stream()))
        for p, r, l in zip(ps, ref, lrs):
            worst = max(worst, ((p - r.detach()).abs() / (r.detach().abs() + l)).max().item())
    # |dp| relative to |p| + lr (one update is O(lr)) after 8 steps.  Measured on B200: 1.1e-6 against torch's foreach
    # implementation (every intermediate rounded to fp32), parameters 1 - 3 ulp apart against its single-kernel one.
    print(f"cp_adam_step vs torch.optim.Adam(fused={fused}): worst |dp| / (|p| + lr) = {worst:.3g}")
    assert worst < 1e-5, worst
    assert counters.tolist() == [8, 8]


def test_adam_step_matches_the_oracle():
    """cp_step_prologue + cp_adam_step against the CPU oracle of the optimiser end of the step: oracle.model.adam_update
    (the rule train.py:72-73's torch.optim.Adam applies; pinned by the reference's own training steps in
    tests/golden/model.npz) on dL/dW + the gradient of reg * ||W||_2 (models.py:225-228) taken by autograd on the CPU."""
    from contrastiveprosthetics_b200 import _lib
    from oracle import model as OM
    L = _lib.lib()
    g = torch.Generator().manual_seed(11)
    shapes, regs, lrs = [(64, 1, 3, 3), (64,), (512, 512), (16, 41)], [1e-3, 0.0, 1e-3, 0.2], [1e-3, 1e-3, 1e-3, 1e-2]
    lr_index, norm_index = [0, 0, 0, 1], [0, -1, 1, 2]
    cpu = [torch.randn(s, generator=g) * 0.1 for s in shapes]
    ps = [p.clone().cuda() for p in cpu]
    m_o, v_o = [torch.zeros_like(p) for p in cpu], [torch.zeros_like(p) for p in cpu]
    n = len(ps)
    offs, off = [], 0
    for p in ps:
        offs.append(off)
        off += (p.numel() + 127) // 128 * 128
    grads, m, v = (torch.zeros(off, device="cuda") for _ in range(3))
    lr = torch.tensor([1e-3, 1e-2], dtype=torch.float64, device="cuda")
    reg_list = [i for i in range(n) if norm_index[i] >= 0]
    norms = torch.zeros(len(reg_list), device="cuda")
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")
    ws = torch.zeros(L.cp_step_workspace_bytes(len(reg_list)), dtype=torch.uint8, device="cuda")
    A = lambda ct, xs: (ct * len(xs))(*xs)                                                # noqa: E731
    worst = 0.0
    for step in range(5):
        gs = [torch.randn(s, generator=g) * (0.1 ** step) for s in shapes]
        for i, (p, gg) in enumerate(zip(cpu, gs)):
            total = gg.clone()
            if norm_index[i] >= 0:                       # d(reg * ||W||)/dW by autograd, like loss.backward() does it
                w = p.clone().requires_grad_(True)
                (torch.norm(w) * regs[i]).backward()
                total = total + w.grad
            OM.adam_update(p, total, m_o[i], v_o[i], step + 1, lrs[i])
        for o, gg, p in zip(offs, gs, ps):
            grads[o:o + p.numel()] = gg.reshape(-1).cuda()
        _lib.check(L.cp_step_prologue(A(ctypes.c_void_p, [ps[i].data_ptr() for i in reg_list]),
                                      A(ctypes.c_int64, [ps[i].numel() for i in reg_list]), len(reg_list), _lib.ptr(norms),
                                      _lib.ptr(counters), 2, _lib.ptr(ws), ws.numel(), _lib.stream()))
        _lib.check(L.cp_adam_step(A(ctypes.c_void_p, [p.data_ptr() for p in ps]), A(ctypes.c_int64, [p.numel() for p in ps]),
                                  A(ctypes.c_int64, offs), n, _lib.ptr(grads), _lib.ptr(m), _lib.ptr(v), _lib.ptr(lr),
                                  A(ctypes.c_int32, lr_index), A(ctypes.c_float, regs), A(ctypes.c_int32, norm_index),
                                  _lib.ptr(norms), _lib.ptr(counters[1:2]), 0.9, 0.999, 1e-8, _lib.stream()))
        for p, r, l in zip(ps, cpu, lrs):
            worst = max(worst, ((p.cpu() - r).abs() / (r.abs() + l)).max().item())
    print(f"cp_adam_step vs oracle adam_update: worst |dp| / (|p| + lr) = {worst:.3g}")
    assert worst < 1e-5, worst


def _autograd_steps(model, batches, fused=True):
    opts = [torch.optim.Adam(model.emg_net.parameters(), lr=PARAMS['lr_emg'], fused=fused),
            torch.optim.Adam(model.glove_net.parameters(), lr=PARAMS['lr_glove'], fused=fused)]
    label = torch.arange(41, device="cuda").repeat(batches[0].shape[0])
    out = []
    for EMG in batches:
        lg = model.forward(EMG, None, label)
        loss = model.loss(lg, label)
        total = loss + model.l2()
        for o in opts:
            o.zero_grad(set_to_none=True)
        total.backward()
        for o in opts:
            o.step()
        out.append((loss.item(), lg.ncor.clone()))
    return out


@pytest.mark.parametrize("adabn", [True, False])
def test_lean_step_equals_autograd_step(adabn):
    """Against the reference's own wiring: autograd + torch.optim.Adam with its default (foreach) implementation
    (train.py:72-73).  Adam's first updates are lr * g / (|g| + eps): an element whose gradient sits at rounding level
    takes a different step under ANY change of rounding, and six steps amplify that -- torch's own two implementations
    (fused=True vs default) end 6e-5 apart in the loss and 3e-3 in single weights on these batches
    (scripts/diag_lean_vs_autograd.py), while the lean step reproduces the default one's losses digit for digit.  So:
    losses to 2e-6, and all but a sliver (0.5 %) of the parameters to 1e-6."""
    from contrastiveprosthetics_b200.step import LeanTrainStep
    batches = _batches(6)
    ref_model = _model(adabn=adabn)
    ref = _autograd_steps(ref_model, batches, fused=False)
    model = _model(adabn=adabn)
    lean = LeanTrainStep(model, PARAMS['lr_emg'], PARAMS['lr_glove'])
    got = []
    for EMG in batches:
        loss, ncor = lean(EMG)
        got.append((loss.item(), ncor.clone()))
    assert got[0][0] == ref[0][0] and torch.equal(got[0][1], ref[0][1])        # same kernels before the first update
    for (l, n), (lr, nr) in zip(got, ref):
        assert abs(l - lr) <= 2e-6 * abs(lr), (l, lr)
    sd, sd_ref = model.state_dict(), ref_model.state_dict()
    off, total = 0, 0
    for k in sd_ref:
        if sd_ref[k].dtype.is_floating_point:
            off += int(((sd[k] - sd_ref[k]).abs() > 1e-6).sum())
            total += sd[k].numel()
        else:
            assert torch.equal(sd[k], sd_ref[k]), k
    print(f"lean vs autograd + Adam (adabn={adabn}): {off} of {total} parameters differ by more than 1e-6")
    assert off <= 5e-3 * total, (off, total)           # measured: 1,824 of 2,027,617 (AdaBN), 0 (stock BN)


def test_lean_gradients_are_the_autograd_gradients():
    """The bucket holds dL/dW of the contrastive loss (bit for bit what autograd returns); the regulariser's gradient
    is added inside cp_adam_step."""
    from contrastiveprosthetics_b200.step import LeanTrainStep
    EMG = _batches(1)[0]
    label = torch.arange(41, device="cuda").repeat(8)
    ref_model = _model()
    ref_model.loss(ref_model.forward(EMG, None, label), label).backward()
    model = _model()
    lean = LeanTrainStep(model, 0.0, 0.0)
    lean(EMG)
    for (name, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        if name == "logit_scale":
            continue
        if q.grad is None:                         # glove_net.last: regularised, not in the forward
            assert not p.grad.any(), name
        else:
            assert torch.equal(p.grad, q.grad), name


@pytest.mark.parametrize("adabn", [True, False])
def test_lean_graph_is_bit_identical_to_lean_eager(adabn):
    """(Without dropout: an eager step mixes its host-side call count into the Philox key, a captured graph only the
    device-side step counter -- same as the autograd step, tests/test_gpu_graph.py.)"""
    from contrastiveprosthetics_b200.graph import GraphedTrainStep
    from contrastiveprosthetics_b200.step import LeanTrainStep
    batches = _batches(6)
    model = _model(adabn=adabn)
    lean = LeanTrainStep(model, PARAMS['lr_emg'], PARAMS['lr_glove'])
    eager = [(l.item(), n.clone()) for l, n in (lean(EMG) for EMG in batches)]
    model2 = _model(adabn=adabn)
    opts = [torch.optim.Adam(model2.emg_net.parameters(), lr=PARAMS['lr_emg']),
            torch.optim.Adam(model2.glove_net.parameters(), lr=PARAMS['lr_glove'])]
    step = GraphedTrainStep(model2, opts, batches[0], lean=True)
    for EMG, (l_ref, n_ref) in zip(batches, eager):
        loss, ncor = step(EMG)
        assert loss.item() == l_ref
        assert torch.equal(ncor, n_ref)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, model.state_dict()[k]), k
    assert step.lean.counters.tolist() == [6, 6]


def test_lean_lr_is_read_at_run_time():
    """A scheduler's new lr reaches the captured graph (lr is a device array); every replay draws a fresh dropout mask."""
    from contrastiveprosthetics_b200.graph import GraphedTrainStep
    model = _model(dp=0.5)
    opts = [torch.optim.Adam(model.emg_net.parameters(), lr=1e-3), torch.optim.Adam(model.glove_net.parameters(), lr=1e-3)]
    EMG = _batches(1)[0]
    step = GraphedTrainStep(model, opts, EMG, lean=True)
    step.lean.set_lr(0.0, 0.0)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    losses = [step(EMG)[0].item() for _ in range(4)]
    assert len(set(losses)) == 4, losses            # frozen weights: only the mask can change the loss
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    step.lean.set_lr(1e-3, 1e-3)
    step(EMG)
    assert not torch.equal(model.state_dict()["emg_net.last.0.weight"], before["emg_net.last.0.weight"])
