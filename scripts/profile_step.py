"""torch.profiler view of one C1 step: every CUDA kernel (ours + torch glue) by total device time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from climategan_b200 import ops
from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_painter_opts
dev = torch.device("cuda:0")
B, S = int(os.environ.get("B", "16")), 640
torch.manual_seed(0)
G = OmniGenerator(default_painter_opts(), latent_shape=S, storage_dtype=torch.bfloat16).to(dev).train()
params = [p for p in G.painter.parameters() if p.requires_grad]
x = torch.rand(B, 3, S, S, device=dev) * 2 - 1
m = (torch.rand(B, 1, S, S, device=dev) > 0.5).float()
t = torch.rand(B, 3, S, S, device=dev) * 2 - 1
def step():
    for p in params: p.grad = None
    loss = ops.l1_loss(G.paint(m, x), t); loss.backward(); return loss
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"), reverse=True)
tot = sum(r[0] for r in rows)
print(f"total CUDA kernel time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for tme, cnt, key in rows[:45]:
    print(f"{tme/1e3:9.3f} ms {cnt:5d}x  {key[:110]}")
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue time {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
