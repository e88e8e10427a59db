"""Diagnostic: error metrics of the bf16 path vs goldens / oracle (prints, no asserts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from climategan_b200 import _lib, ops
from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_painter_opts
from tests.helpers import load_golden, rel_max, rel_l2, cosine
from oracle import painter_oracle as po
cuda = torch.device("cuda:0")

def q(x, dt): return x.to(dt).float()
for dtype in (torch.float32, torch.bfloat16):
    for act in (_lib.ACT_NONE, _lib.ACT_LRELU):
        for c, h, w in [(20, 12, 12), (40, 6, 10), (128, 4, 4)]:
            torch.manual_seed(7 * c + h)
            n = 2
            x = q(torch.randn(n, c, h, w) * 1.5 + 0.3, dtype); seg = q(torch.rand(n, 3, h, w) * 2 - 1, dtype)
            sd = {"p.mlp_shared.0.weight": torch.randn(128, 3, 3, 3) * 0.3, "p.mlp_shared.0.bias": torch.randn(128) * 0.1,
                  "p.mlp_gamma.weight": torch.randn(c, 128, 3, 3) * 0.03, "p.mlp_gamma.bias": torch.randn(c) * 0.1,
                  "p.mlp_beta.weight": torch.randn(c, 128, 3, 3) * 0.03, "p.mlp_beta.bias": torch.randn(c) * 0.1}
            sd = {k: q(v, dtype) if k.endswith("weight") else v for k, v in sd.items()}
            sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
            xr = x.double().requires_grad_(True)
            out_r = po.spade(sdr, "p", xr, seg.double())
            if act == _lib.ACT_LRELU: out_r = F.leaky_relu(out_r, 0.2)
            go = q(torch.randn_like(out_r).float(), dtype); out_r.backward(go.double())
            sdg = {k: v.to(cuda).requires_grad_(True) for k, v in sd.items()}
            xs = ops.to_storage(x.to(cuda), dtype).requires_grad_(True); segs = ops.to_storage(seg.to(cuda), dtype)
            mean, rstd = ops.instnorm_stats(xs)
            out = ops.spade(xs, mean, rstd, segs, sdg["p.mlp_shared.0.weight"], sdg["p.mlp_shared.0.bias"], sdg["p.mlp_gamma.weight"], sdg["p.mlp_gamma.bias"], sdg["p.mlp_beta.weight"], sdg["p.mlp_beta.bias"], act, 0.2)
            o = ops.from_storage(out, c); o.backward(go.to(cuda))
            gx = ops.from_storage(xs.grad, c)
            print(f"spade {dtype} act{act} c{c}: out relmax {rel_max(o,out_r):.2e} gx relmax {rel_max(gx,xr.grad):.2e} l2 {rel_l2(gx,xr.grad):.2e} cos {cosine(gx,xr.grad):.5f}", end=" ")
            print(" ".join(f"{k.split('.')[1][:6]}{'W' if k.endswith('weight') else 'b'}:{rel_l2(sdg[k].grad, sdr[k].grad):.1e}" for k in sd))

meta, g, sd, (x, m, t) = load_golden()
for dtype in (torch.float32, torch.bfloat16):
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
    G = OmniGenerator(opts, latent_shape=meta["size"], storage_dtype=dtype); G.painter.load_state_dict(sd); G = G.to(cuda)
    out = G.paint(m.to(cuda), x.to(cuda)); loss = ops.l1_loss(out, t.to(cuda)); loss.backward()
    print(dtype, "out relmax", rel_max(out, torch.from_numpy(g["out"])), "l2", rel_l2(out, torch.from_numpy(g["out"])), "loss", float(loss), float(g["loss"]))
    params = dict(G.painter.named_parameters())
    for k, v in g.items():
        if k.startswith("grad::"):
            gm = params[k[6:]].grad; gr = torch.from_numpy(v)
            print("   ", k[6:], f"relmax {rel_max(gm,gr):.2e} l2 {rel_l2(gm,gr):.2e} cos {cosine(gm,gr):.5f}")
    nr = dict(zip(meta["grad_keys"], g["grad_norms"]))
    worst = sorted(((abs(float(params[k].grad.norm()) - nr[k]) / nr[k], k) for k in meta["grad_keys"] if nr[k] > 1e-6), reverse=True)[:5]
    print("   worst grad-norm rel errs", worst)
