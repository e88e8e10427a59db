"""Two forward calls of an AdvEnt discriminator before one backward (the D step sees the r and the s batch): product (fp32) vs oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from climategan_b200.discriminator import get_fc_discriminator, fc_discriminator_forward
from climategan_b200 import ops
from oracle import full_step_oracle as fo
from oracle.painter_oracle import SNState

dev = torch.device("cuda:0")
torch.manual_seed(0)
for ncls, hw in ((2, 64), (11, 32)):
    net = get_fc_discriminator(num_classes=ncls, use_norm=True)
    sd = {"D." + k: v.detach().clone() for k, v in net.state_dict().items()}
    pa = torch.softmax(torch.randn(2, ncls, hw, hw), 1)
    pb = torch.softmax(torch.randn(2, ncls, hw, hw), 1)
    for flip in (False, True):
        # oracle
        osd = {k: v.clone().requires_grad_(flip or not k.endswith(("_u", "_v"))) for k, v in sd.items()}
        sn = SNState(osd)
        lo = fo.advent(pa, 1, osd, sn, "D", None, wgan=False) + fo.advent(pb, 0, osd, sn, "D", None, wgan=False)
        lo.backward()
        # product
        net2 = get_fc_discriminator(num_classes=ncls, use_norm=True)
        net2.load_state_dict({k[2:]: v for k, v in sd.items()})
        net2 = net2.to(dev)
        for n, p in net2.named_parameters():
            p.requires_grad_(flip or not n.endswith(("_u", "_v")))
        D = lambda t: fc_discriminator_forward(net2, t, torch.float32)
        lp = ops.const_target_loss(D(ops.prob_2_entropy(pa.to(dev))), ops.LOSS_BCE_LOGITS, 1.0) + \
             ops.const_target_loss(D(ops.prob_2_entropy(pb.to(dev))), ops.LOSS_BCE_LOGITS, 0.0)
        lp.backward()
        print(ncls, "flip", flip, "loss", float(lo), float(lp))
        for n, p in net2.named_parameters():
            go = osd["D." + n].grad
            if go is None or p.grad is None:
                print("   ", n, "grad None", go is None, p.grad is None); continue
            e = float((p.grad.cpu() - go).abs().max() / go.abs().max().clamp_min(1e-30))
            if e > 1e-4: print("   ", n, "rel", e, float(go.norm()), float(p.grad.norm()))
        for n, p in net2.named_parameters():
            if n.endswith(("_u", "_v")):
                e = float((p.detach().cpu() - osd["D." + n].detach()).abs().max())
                if e > 1e-5: print("    state", n, e)
