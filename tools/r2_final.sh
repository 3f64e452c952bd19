#!/bin/bash
# round-2 acceptance run on one GPU: the whole GPU test-suite, smoke, both bench arms as the driver launches them
O=gpurun_out/$1; mkdir -p $O
( time timeout 2400 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee $O/tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/smoke.txt
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2>&1 | grep real
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real
tail -3 $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
r=json.loads([l for l in open("$O/bench_ref.json") if l.startswith("{")][-1])
print("value %.0f e2e %.0f frac %.3f kernel_ms %.4f ref %.2f ratio %.0f e2e_ratio %.0f same_config %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], r["value"], d["value"]/r["value"], d["e2e"]["value"]/r["value"], d["config"]==r["config"]))
print("sustained", d["sustained"]["value"], d["sustained"]["seconds"], d["sustained"]["power_w_max"])
for k,v in (d.get("configs") or {}).items():
    print(k, "ms %.3f Mrays/s %.0f e2e %.0f" % (v["ms_per_frame"], v["value"], v["e2e"]["value"]), v.get("roofline",{}).get("frac"), v.get("executed",{}).get("exact_tests_per_ray_query"), v.get("cpu_baseline",{}).get("value"))
s=d["strong_config5"]; print("strong", s["ms_frame_solo_rank0"], s["ms_frame_split"], s["pixels_equal"], s["e2e"]["ms_per_frame"])
print(d["cpu_baseline"])
PY
