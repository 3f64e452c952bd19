#!/bin/bash
# bench.py under torchrun at N ranks (the driver's launch line).  Usage: r2_scale.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift; shift
O=gpurun_out/$TAG; mkdir -p $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2>&1 | grep real
grep -v "OMP_NUM\|^\*\*\*\|^$" $O/bench_n$N.err | tail -8
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n$N.json") if l.startswith("{")][-1])
    print("N=%d value %.0f e2e %.0f ms/step %.3f gather_parity %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("gather_parity")))
    print(json.dumps(d.get("gather_parity_detail"))[:600])
    s=d.get("strong_config5"); 
    if s: print("strong: solo %.1f split %.1f speedup %.2f equal %s e2e %.1f ms" % (s["ms_frame_solo_rank0"], s["ms_frame_split"], s["speedup"], s["pixels_equal"], s["e2e"]["ms_per_frame"]))
    print(d.get("sustained"))
except Exception as e:
    print("FAILED", e)
PY
