#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,launch__grid_size,launch__shared_mem_per_block_dynamic,sm__cycles_elapsed.avg"
for v in 0 1; do
ONLY=w_r1,w_r1b,w_r3 CGB_TC2_WGRAD=$v timeout 300 ncu --metrics $M --clock-control none -k regex:wgrad_tc -s 3 -c 6 --csv --log-file gpurun_out/g12_wgrad_pair$v.csv python scripts/exp/tc2_check.py save > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.DictReader(l for l in open("gpurun_out/g12_wgrad_pair$v.csv") if l.startswith('"')))
agg = {}
for r in rows:
    agg.setdefault((r["ID"], r["Kernel Name"][:20], r.get("Grid Size") or ""), {})[r["Metric Name"]] = r["Metric Value"]
for k, m in agg.items():
    print("pair=$v", k[1], {a.split("__")[-1][:28]: b for a, b in m.items()})
PY
done
