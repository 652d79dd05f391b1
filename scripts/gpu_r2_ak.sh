#!/bin/bash
# Round 2, GPU call AK: ncu launch list of the default bench command past its set-up launches (the timed graph replays),
# ncu --set full of the batched launch summarised on the box.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 560 --launch-count 2500 --csv --log-file gpurun_out/r2ak_launches_bench.csv python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --batch 0 --min-region-ms 0.05 > gpurun_out/r2ak_launches_bench.json 2> gpurun_out/r2ak_launches.err; echo "launch list rc=$?"; tail -2 gpurun_out/r2ak_launches.err
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2ak_launches_bench.csv')))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        k=d['Kernel Name'][:70]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{v[0]:5d} {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]/1e3:.2f} us  {k}")
PY
timeout 600 ncu --set full --clock-control none -k regex:spmm_edgelist_kernel -s 1 -c 1 -f -o /tmp/r2ak_full_batched python scripts/batch_probe.py > gpurun_out/r2ak_full_batched.log 2>&1; echo "ncu full batched rc=$?"
python scripts/ncu_summary.py gpurun_out/r2ak_ncu_batched.md "ncu --set full of ONE batched launch (sx_spmm_device_batch_f64: 20 operand triples of nasa4704 N=16, grid 147 x 20)" /tmp/r2ak_full_batched.ncu-rep; cat gpurun_out/r2ak_ncu_batched.md | tail -5
