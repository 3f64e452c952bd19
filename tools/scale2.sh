#!/bin/bash
# N=2 validation of the multi-GPU bench modes (run under gpurun --gpus 2)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -3
$TR bench.py --gpus 2 --steps 10 --warmup 3 --gather store > gpurun_out/s2_store.json 2> gpurun_out/s2_store.err
$TR bench.py --gpus 2 --steps 10 --warmup 3 --gather copy > gpurun_out/s2_copy.json 2> gpurun_out/s2_copy.err
C5="--config 5 --spp 4 --frames-per-step 1 --steps 4 --warmup 3 --count-one --count-fast --cpu-fraction 0.004"
python bench.py --gpus 1 $C5 > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err
$TR bench.py --gpus 2 $C5 --tile-split > gpurun_out/c5_n2.json 2> gpurun_out/c5_n2.err
tail -n 3 gpurun_out/s2_store.err gpurun_out/s2_copy.err gpurun_out/c5_n1.err gpurun_out/c5_n2.err
python - <<PY
import json
for n in ("s2_store","s2_copy","c5_n1","c5_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%n))
        print(n, "value %.1f ms/step %.3f fps %.1f e2e %.1f scaling %s" % (d["value"], d["ms_per_step"], d["frames_per_s"], d["e2e"]["value"], d["scaling"]))
    except Exception as e:
        print(n, "FAILED", e)
PY
