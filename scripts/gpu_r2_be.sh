#!/bin/bash
# Round 2, GPU call BE: ncu DRAM traffic per launch of the edge-list kernel after the row-aligned streams (feeds profiles/traffic.json).
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
prof() { tag=$1; shift; ncu --metrics $M --clock-control none -k regex:"spmm_edgelist" -s 12 -c 4 --csv --log-file gpurun_out/r2be_ncu_$tag.csv python bench.py --configs none --no-cpu-baseline --no-pipelined-e2e --batch 0 --no-graph --steps 3 --warmup 3 --min-region-ms 0.01 "$@" > /dev/null 2> gpurun_out/r2be_ncu_$tag.err; echo "ncu $tag rc=$?"; }
prof nasa4704_n16_f64
for n in 8 16 32 64; do prof pcrystk02_n${n}_f32 --workload pcrystk02 --ncols $n; done
python - <<'PY'
import csv,glob,json,os
out={}
for f in sorted(glob.glob('gpurun_out/r2be_ncu_*.csv')):
    tag=os.path.basename(f)[len('r2be_ncu_'):-4]
    rows=list(csv.reader(open(f)))
    hdr=None; per={}
    for r in rows:
        if 'Kernel Name' in r: hdr=r; continue
        if hdr and len(r)==len(hdr):
            d=dict(zip(hdr,r)); i=d['ID']; per.setdefault(i,{})
            try: per[i][d['Metric Name']]=(float(d['Metric Value'].replace(',','')), d['Metric Unit'])
            except: pass
    def tobytes(v,u): return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    vals=[tobytes(*p['dram__bytes_read.sum'])+tobytes(*p['dram__bytes_write.sum']) for p in per.values() if 'dram__bytes_read.sum' in p]
    times=[p['gpu__time_duration.sum'][0] for p in per.values() if 'gpu__time_duration.sum' in p]
    if vals: out[tag]={'traffic':int(sorted(vals)[len(vals)//2]), 'n':len(vals), 'time':sorted(times)[len(times)//2] if times else None}
print(json.dumps(out,indent=1))
json.dump(out,open('gpurun_out/r2be_traffic.json','w'),indent=1)
PY
