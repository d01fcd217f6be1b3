"""CUDA-graph captured training step (graph.GraphedTrainStep) == the eager step, launch for launch."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5}


def _setup(dp, adabn=True):
    from contrastiveprosthetics_b200.models import Model
    p = dict(PARAMS, dp_emg=dp)
    torch.manual_seed(42)
    model = Model(p, adabn=adabn, device="cuda")
    model.set_train()
    opts = [torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, capturable=True),
            torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, capturable=True)]
    return model, opts


def _batches(n, B=8):
    g = torch.Generator().manual_seed(1)
    return [(torch.randn(B, 41, 1, 1, 12, generator=g) + 0.5 * torch.randn(1, 41, 1, 1, 12, generator=g)).cuda()
            for _ in range(n)]


@pytest.mark.parametrize("adabn", [True, False])
def test_graphed_step_is_bit_identical_to_eager(adabn):
    from contrastiveprosthetics_b200.graph import GraphedTrainStep
    batches = _batches(6)
    label = torch.arange(41, device="cuda").repeat(8)
    # eager
    model, opts = _setup(0.0, adabn)
    eager = []
    for EMG in batches:
        lg = model.forward(EMG, None, label)
        loss = model.loss(lg, label)
        total = loss + model.l2()
        for o in opts:
            o.zero_grad(set_to_none=True)
        total.backward()
        for o in opts:
            o.step()
        eager.append((loss.item(), lg.ncor.clone()))
    sd_eager = {k: v.clone() for k, v in model.state_dict().items()}
    # graph
    model2, opts2 = _setup(0.0, adabn)
    step = GraphedTrainStep(model2, opts2, batches[0])
    for EMG, (l_ref, n_ref) in zip(batches, eager):
        loss, ncor = step(EMG)
        assert loss.item() == l_ref
        assert torch.equal(ncor, n_ref)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, sd_eager[k]), k


def test_graphed_step_draws_fresh_dropout_masks():
    from contrastiveprosthetics_b200.graph import GraphedTrainStep
    model, opts = _setup(0.5)
    for o in opts:
        for g in o.param_groups:
            g["lr"] = 0.0                           # frozen weights: only the mask can change the loss
    EMG = _batches(1)[0]
    step = GraphedTrainStep(model, opts, EMG)
    losses = [step(EMG)[0].item() for _ in range(4)]
    assert len(set(losses)) == 4, losses
