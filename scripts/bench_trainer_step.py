"""Time the full painter train step (update_G + update_D, VGG + GAN + featmatch, ExtraAdam) at 640x640."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climategan_b200.trainer import Trainer
from climategan_b200.utils import default_painter_opts
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "8")); S = 640
opts = default_painter_opts()
torch.manual_seed(0)
t = Trainer(opts, device=dev, storage_dtype=torch.bfloat16).setup(input_shape=(S, S))
x = torch.rand(B, 3, S, S, device=dev) * 2 - 1
m = (torch.rand(B, 1, S, S, device=dev) > 0.5).float()
batch = {"rf": {"data": {"x": x, "m": m}}}
def step():
    t.update_G(batch); t.update_D(batch); t.logger.global_step += 1
for _ in range(2): step()
torch.cuda.synchronize()
from climategan_b200 import _lib
lib = _lib.lib(); lib.cgb_prof_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); n = 3
for _ in range(n): step()
e1.record(); torch.cuda.synchronize(); lib.cgb_prof_enable(0)
ms = e0.elapsed_time(e1) / n
print(f"painter G+D step B={B}: {ms:.1f} ms/step  {B/ms*1e3:.1f} img/s  losses {t.losses_to_host()}  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
import ctypes as C
buf = C.create_string_buffer(1 << 20); lib.cgb_prof_dump(buf, len(buf))
rows = []
for ln in buf.value.decode().strip().splitlines():
    f = ln.split(); which, tc, n_, hi, wi, ci, ho, wo, co, kh, kw, stride, dil, count = map(int, f[:14]); tot = float(f[14])
    rows.append((tot / n, count // n, ["fwd","dgrad","wgrad"][which], "tc" if tc else "simt", ci, co, kh, stride, hi, wi))
rows.sort(reverse=True)
for r in rows[:14]: print("%8.3f ms/step %3dx %s[%s] %d->%d k%d s%d @%dx%d" % r)
