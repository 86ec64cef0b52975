"""GPU (-m gpu): a whole training step captured in a CUDA graph (utils/graph.py, FusedAdamW(capturable=True)) must
train exactly like the eager step: same losses and same parameters after several replays, with a learning-rate
schedule stepping between replays."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from ucf_vit_b200.simple.arch import VIT
    torch.manual_seed(0)
    return VIT(img_size=[64, 64], patch_size=8, in_chans=3, num_classes=5, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4,
               class_token=True, twoD=True, default_vars=["r", "g", "b"]).cuda().train()


def test_graphed_step_matches_eager_step():
    from ucf_vit_b200.utils.graph import GraphedTrainStep
    from ucf_vit_b200.utils.misc import configure_optimizer, configure_scheduler
    m_e = _model()
    m_g = copy.deepcopy(m_e)
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(8, 3, 64, 64, generator=g).cuda() for _ in range(6)]
    ys = [torch.randint(0, 5, (8,), generator=g).cuda() for _ in range(6)]
    lossf = torch.nn.CrossEntropyLoss()

    def make_step(model, opt):
        def step(x, y):
            loss = lossf(model(x, ["r", "g", "b"]).float(), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        return step

    opt_e = configure_optimizer(m_e, 1e-3, 0.9, 0.95, 1e-5, fused="ucf")
    sch_e = configure_scheduler(opt_e, 2, 10, 1e-5, 1e-6)
    opt_g = configure_optimizer(m_g, 1e-3, 0.9, 0.95, 1e-5, fused="ucf_capturable")
    sch_g = configure_scheduler(opt_g, 2, 10, 1e-5, 1e-6)
    step_e = make_step(m_e, opt_e)
    # the capture's warm-up steps must not train the model: snapshot and restore around construction
    state = copy.deepcopy(m_g.state_dict())
    graphed = GraphedTrainStep(make_step(m_g, opt_g), (xs[0], ys[0]), warmup=2)
    m_g.load_state_dict(state)
    for st in opt_g.state.values():
        st["exp_avg"].zero_()
        st["exp_avg_sq"].zero_()
    opt_g._step_dev.zero_()
    losses_e, losses_g = [], []
    for x, y in zip(xs, ys):
        losses_e.append(step_e(x, y).item())
        sch_e.step()
        losses_g.append(graphed(x, y).item())
        sch_g.step()
    assert graphed.replays == 6
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (losses_e, losses_g)
    assert float(opt_g._step_dev) == 6.0
    assert abs(float(opt_g.param_groups[0]["lr"]) - float(opt_e.param_groups[0]["lr"])) < 1e-9
    # Adam's first updates are sign-like (|update| ~ lr whatever the gradient's size), so elements whose gradient is
    # within the bf16 / split-K summation noise may move the other way: compare the UPDATE VECTORS, not elements
    init = dict(state)
    for (k, a), (_, b) in zip(m_e.named_parameters(), m_g.named_parameters()):
        da, db = (a.detach() - init[k]).double().flatten(), (b.detach() - init[k]).double().flatten()
        if da.norm() == 0:
            continue
        cos = torch.nn.functional.cosine_similarity(da, db, dim=0).item()
        assert cos >= 0.95, f"{k}: update cosine {cos:.4f}"
        if a.dim() >= 2:      # (biases: the key bias' true gradient is zero, its Adam update is +-lr noise)
            assert (a - b).norm() <= 2e-2 * a.norm() + 1e-6, k


def test_adamw_device_hyperparameters_match_host_form():
    from ucf_vit_b200 import ops
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(n, generator=g).cuda() for n in (5000, 17, 4096)]
    gs = [torch.randn(p.shape, generator=g).cuda() for p in ps]
    a = [p.clone() for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    b = [p.clone() for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    lr_dev = torch.tensor(3e-3, device="cuda")
    step_dev = torch.zeros(1, device="cuda")
    for step in range(1, 8):
        ops.adamw_multi(a[0], gs, a[1], a[2], lr=3e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=1e-2, step=step)
        step_dev.add_(1.0)
        ops.adamw_multi(b[0], gs, b[1], b[2], lr=lr_dev, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=1e-2, step=step_dev)
    for x, y in zip(a[0] + a[1] + a[2], b[0] + b[1] + b[2]):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-7)
