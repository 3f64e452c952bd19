#!/bin/bash
# Every launch of a short bench run with its device time (profiles/r2_bench_launches.csv): the trace kernel's SHARE of a
# step must agree with bench.py's CUDA-event figure.
O=gpurun_out/$1; mkdir -p $O
CMD="python bench.py --steps 5 --warmup 3 --sustained-seconds 0"
$CMD > $O/bench_short.json 2> $O/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv $CMD > $O/ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/launches.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
col={h:i for i,h in enumerate(rows[hdr])}
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    try: v=float(r[col["Metric Value"]].replace(",",""))
    except Exception: continue
    u=r[col["Metric Unit"]]
    v = v*1e-3 if u in ("ns","nsecond") else v*1e3 if u in ("ms","msecond") else v if u in ("us","usecond") else v*1e6
    k=r[col["Kernel Name"]][:100]
    agg[k][0]+=1; agg[k][1]+=v
with open("$O/launches_summary.txt","w") as f:
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        line="%-100s launches %6d total %12.1f us avg %10.2f us" % (k,n,t,t/n)
        print(line); f.write(line+"\n")
PY
gzip -f $O/launches.csv
