#!/bin/bash
# BASELINE.json configs[4] as specified: 3840x2160, 1024 spheres, depth 8, 64 spp; one frame cut into super-tiles over 1 and 8 GPUs
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
C5="--config 5 --frames-per-step 1 --steps 3 --warmup 3 --count-one --count-fast --cpu-fraction 0.001 --tile-split"
python bench.py --gpus 1 $C5 > gpurun_out/c5f_n1.json 2> gpurun_out/c5f_n1.err
$TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 $C5 > gpurun_out/c5f_n8.json 2> gpurun_out/c5f_n8.err
python - <<PY
import json
for n in ("c5f_n1","c5f_n8"):
    try:
        txt=open("gpurun_out/%s.json"%n).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "value %.1f ms/step %.3f fps %.3f e2e %.1f scaling %s rays/step %.0f cpu %s" % (d["value"], d["ms_per_step"], d["frames_per_s"], d["e2e"]["value"], d["scaling"], d["config"]["rays_per_step"], d.get("cpu_baseline")))
    except Exception as e:
        print(n, "FAILED", e)
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/c5f_n8.err | tail -5
