"""Print per-tensor parity of a golden case (product on GPU vs fp32 CPU oracle)."""
import sys
import torch
sys.path.insert(0, ".")
from tests import _cases as C

name = sys.argv[1] if len(sys.argv) > 1 else "unetr_3d_var2"
cfg, shapes, arrays, sd = C.load(name)
inp = C.inputs(cfg, arrays)
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
o_out, o_loss = C.run_oracle(cfg, sdg, inp)
o_loss.backward()
model = C.build_product(cfg)
model.load_state_dict(sd, strict=True)
model = model.cuda().train(cfg.get("train", True))
p_out, p_loss = C.run_product(cfg, model, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()})
p_loss.backward()
torch.cuda.synchronize()
print("loss", p_loss.item(), o_loss.item())
for k, ref in o_out.items():
    got = p_out[k].detach().float().cpu()
    print(f"out {k:24s} rel {((got-ref.detach()).norm()/ref.detach().norm()).item():.3e}")
named = dict(model.named_parameters())
gmax = max(v.grad.norm().item() for v in sdg.values() if v.grad is not None)
for k, v in sdg.items():
    if v.grad is None or k not in named or named[k].grad is None:
        continue
    g = named[k].grad.detach().float().cpu()
    rel = ((g - v.grad).norm() / (v.grad.norm() + 1e-30)).item()
    cos = torch.nn.functional.cosine_similarity(g.double().flatten(), v.grad.double().flatten(), dim=0).item()
    print(f"grad {k:44s} norm {v.grad.norm().item():.3e} ({v.grad.norm().item()/gmax:.1e} of max) rel {rel:.3e} cos {cos:.5f}")
