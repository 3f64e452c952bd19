#!/bin/bash
# config 5 (3840x2160, 1024 spheres, depth 8) at 4 spp: ONE frame cut into super-tiles over 1/4/8 GPUs (strong scaling)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
C5="--config 5 --spp 4 --frames-per-step 1 --steps 4 --warmup 3 --count-one --count-fast --cpu-fraction 0.004 --tile-split"
python bench.py --gpus 1 $C5 > gpurun_out/c5t_n1.json 2> gpurun_out/c5t_n1.err
$TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 $C5 > gpurun_out/c5t_n4.json 2> gpurun_out/c5t_n4.err
$TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 $C5 > gpurun_out/c5t_n8.json 2> gpurun_out/c5t_n8.err
python - <<PY
import json
for n in ("c5t_n1","c5t_n4","c5t_n8"):
    try:
        txt=open("gpurun_out/%s.json"%n).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "value %.1f ms/step %.3f fps %.2f e2e %.1f scaling %s" % (d["value"], d["ms_per_step"], d["frames_per_s"], d["e2e"]["value"], d["scaling"]))
    except Exception as e:
        print(n, "FAILED", e)
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/c5t_n8.err | tail -5
