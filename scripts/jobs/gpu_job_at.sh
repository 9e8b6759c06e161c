#!/bin/bash
mkdir -p gpurun_out
for C in 4096 11008; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chol_launches_$C.csv python scripts/chol_ncu.py $C > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/chol_launches_$C.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
rows = rows[1:]
# second half = second call?  the script runs ONE call after warm-up-free setup: take everything
agg = collections.OrderedDict()
for r in rows:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
print("C=$C")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k:70s} n={n:4d} total {t/1e3:8.3f} ms avg {t/n:7.1f} us")
PY
done
