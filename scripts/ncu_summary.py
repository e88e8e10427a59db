"""Summarise an .ncu-rep (read with `ncu -i`) into text: key metrics + top stall sites.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for krow in rows[2:]:
    name = krow[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"## kernel: {name}")
    for h, u, v in zip(hdr, units, krow):
        if any(h == k or h.endswith("." + k) or h.endswith(k) for k in KEYS):
            print(f"  {h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page repeats a 2-line header per kernel; take the first kernel
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"\n## top stall sites ({tot} warp samples)")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    s = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[:2]
    print(f"  {100*s/tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} exec={r[ix['Instructions Executed']]:>10s}  {why[0][1]}")
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("\n## stall reasons (all samples)")
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {h:28s} {100*v/tot:5.1f}%")
