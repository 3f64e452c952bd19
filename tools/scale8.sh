#!/bin/bash
# N=4/8 validation of the frame-dealt multi-GPU bench (run under gpurun --gpus 8)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --gather copy > gpurun_out/s8_copy.json 2> gpurun_out/s8_copy.err
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --gather store > gpurun_out/s8_store.json 2> gpurun_out/s8_store.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --gather copy > gpurun_out/s4_copy.json 2> gpurun_out/s4_copy.err
python - <<PY
import json
for n in ("s4_copy","s8_copy","s8_store"):
    try:
        txt=open("gpurun_out/%s.json"%n).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "value %.1f ms/step %.3f fps %.1f e2e %.1f (%.1f fps) scaling %s lines %d" % (d["value"], d["ms_per_step"], d["frames_per_s"], d["e2e"]["value"], d["e2e"]["frames_per_s"], d["scaling"], len(txt.splitlines())))
    except Exception as e:
        print(n, "FAILED", e)
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/s8_copy.err | tail -5
