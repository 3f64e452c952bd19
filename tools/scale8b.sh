#!/bin/bash
# N=8 / N=4 rerun (NUMA-bound ranks), plus host topology facts for the write-up
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "^CPU\(s\)|NUMA node|Socket|Model name" | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s8b.json 2> gpurun_out/s8b.err
$TR --nproc-per-node 4 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/s4b.json 2> gpurun_out/s4b.err
python - <<PY
import json
for n in ("s4b","s8b"):
    try:
        txt=open("gpurun_out/%s.json"%n).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "value %.1f ms/step %.3f fps %.1f e2e %.1f (%.1f fps) scaling %s lines %d" % (d["value"], d["ms_per_step"], d["frames_per_s"], d["e2e"]["value"], d["e2e"]["frames_per_s"], d["scaling"], len(txt.splitlines())))
    except Exception as e:
        print(n, "FAILED", e)
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/s8b.err | tail -5
