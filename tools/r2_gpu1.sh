#!/bin/bash
# round 2: new GPU tests, bench N=1 (both arms), A/B of the no-fuse variant
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_api_contract.py tests/test_gpu_fullsize_windows.py -m gpu -x -q 2>&1 | tail -25 | tee $O/tests.txt
( time timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err ) 2>&1 | grep real
( time timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real
tail -3 $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print("value %.0f e2e %.0f frac %.3f kernel_ms %.4f sustained %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d.get("sustained")))
for k,v in (d.get("configs") or {}).items():
    print(k, "ms %.3f Mrays/s %.0f e2e %.0f" % (v["ms_per_frame"], v["value"], v["e2e"]["value"]), v.get("roofline",{}).get("frac"), v.get("cpu_baseline",{}).get("value"))
print(d.get("strong_config5"))
print(d.get("cpu_baseline"))
PY
bash tools/r2_ab.sh r2e librtrb_nofuse2.so 2>&1 | grep -v "^config [12] "
