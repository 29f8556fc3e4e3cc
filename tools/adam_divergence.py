"""Why torch.optim.Adam(fused=True) / Adam(capturable=True) and the plain Adam "change the training trajectory" (VERDICT r01, weak 1d).

The same parameters, the same stream of gradients, N steps of Adam with the reference's schedule (lr x 0.997 every 5 steps, PPO.py:52,216-220) in
  plain      : torch.optim.Adam(lr=float)                 -- the reference's optimiser: bias corrections and step size in float64 on the host
  capturable : torch.optim.Adam(lr=tensor, capturable)    -- what the graphed update uses: the same formulas in fp32 on the device
  fused      : torch.optim.Adam(fused=True)               -- one multi-tensor kernel, fp32 scalars on the device
  fp64       : plain Adam on float64 copies                -- the yardstick
Prints, per variant, the distance to fp64 after 1 / 25 / 250 steps relative to the distance travelled; and the per-step relative difference
between plain and capturable step sizes (the whole effect: 1 - beta^t and lr / bias_correction1 rounded to fp32 instead of float64)."""
import json, sys
import torch

dev = "cuda"
torch.manual_seed(0)
shapes = [(264, 460), (264,), (264, 264), (264,), (6, 264), (64, 130)]
base = [torch.randn(s, device=dev) * 0.05 for s in shapes]
N = 250
gen = torch.Generator(device=dev); gen.manual_seed(1)
grads = [[torch.randn(s, device=dev, generator=gen) * (0.01 if len(s) > 1 else 0.1) for s in shapes] for _ in range(N)]


def run(kind):
    dt = torch.float64 if kind == "fp64" else torch.float32
    ps = [b.detach().clone().to(dt).requires_grad_(True) for b in base]
    if kind == "capturable":
        opt = torch.optim.Adam(ps, lr=torch.tensor(0.00014, device=dev), capturable=True)
    elif kind == "fused":
        opt = torch.optim.Adam(ps, lr=0.00014, fused=True)
    else:
        opt = torch.optim.Adam(ps, lr=0.00014)
    snaps = {}
    for t in range(N):
        if t % 5 == 0:
            for g in opt.param_groups:
                g["lr"] *= 0.997
        for p, g in zip(ps, grads[t]):
            p.grad = g.to(dt)
        opt.step()
        if t + 1 in (1, 25, 250):
            snaps[t + 1] = [p.detach().double().clone() for p in ps]
    return snaps


res = {k: run(k) for k in ("fp64", "plain", "capturable", "fused")}
out = {}
for k in ("plain", "capturable", "fused"):
    out[k] = {}
    for t in (1, 25, 250):
        num = sum(float((a - b).abs().max()) for a, b in zip(res[k][t], res["fp64"][t]))
        den = sum(float((b - b0.double()).abs().max()) for b, b0 in zip(res["fp64"][t], base))
        out[k][f"max_abs_err_vs_fp64_over_distance_travelled_step_{t}"] = num / den
out["plain_vs_capturable_step_250_rel"] = sum(float((a - b).abs().max()) for a, b in zip(res["plain"][250], res["capturable"][250])) / \
    sum(float((b - b0.double()).abs().max()) for b, b0 in zip(res["fp64"][250], base))
print(json.dumps(out))
