#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
VLMC_POTRF_V1=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_chol4096_launches_v1_$v.csv python scripts/chol_ncu.py 4096 > /dev/null 2>&1
python - <<PY
import csv, collections, re
rows=list(csv.reader(open('gpurun_out/r02o_chol4096_launches_v1_$v.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    n=re.sub(r'\(.*','',r[kn])[:60]; agg[n][0]+=1; agg[n][1]+=float(r[mv].replace(',',''))
print("VLMC_POTRF_V1=$v")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print(f"  {k:60s} n={v[0]:4d} total {v[1]/1e6:7.3f} ms avg {v[1]/v[0]/1e3:7.1f} us")
PY
done
