"""torch.profiler view of one full train step (update_G + update_D): every CUDA kernel (ours + torch glue) by total device
time, plus the host enqueue time (is the step launch-bound?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts, synth_batch
dev = torch.device("cuda:0")
B, S = int(os.environ.get("B", "8")), 640
torch.manual_seed(0)
opts = full_opts(nblocks=(3, 4, 23, 3), size=S, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3)
t = Trainer(opts, device=dev, storage_dtype=torch.bfloat16).setup(input_shape=(S, S))
mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, B, S, 1).items()}
def step():
    t.update_G(mdb); t.update_D(mdb); t.logger.global_step += 1
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"), reverse=True)
tot = sum(r[0] for r in rows)
print(f"total CUDA kernel time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for tme, cnt, key in rows[:150]:
    print(f"{tme/1e3:9.3f} ms {cnt:5d}x  {key[:120]}")
cpu_rows = sorted(((e.self_cpu_time_total, e.count, e.key) for e in ev if e.self_cpu_time_total > 0), reverse=True)
print(f"--- host side: self CPU time by op (total {sum(r[0] for r in cpu_rows)/1e3:.1f} ms)")
for tme, cnt, key in cpu_rows[:40]:
    print(f"{tme/1e3:9.3f} ms {cnt:5d}x  {key[:100]}")
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue time {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
print(f"max memory allocated {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
