#!/bin/bash
# persistent warp-owned kernel vs one-shot lockstep CTAs on the many-sphere scene, by sample count
O=gpurun_out/$1; mkdir -p $O
for lib in librtrb_w0.so librtrb_w32.so; do
  export RTRB_B200_LIB=$PWD/raytracing_rb_b200/csrc/variants/$lib
  echo "=== $lib" | tee -a $O/warp_sweep.txt
  for spp in 1 2 4 8 16 32; do
    timeout 120 python tools/run_config.py 5 --width 480 --height 270 --spp $spp --frames 3 | tail -1 | tee -a $O/warp_sweep.txt
  done
  timeout 300 python tools/run_config.py 5 --spp 4 --frames 2 | tail -1 | tee -a $O/warp_sweep.txt
done
RTRB_B200_LIB=$PWD/raytracing_rb_b200/csrc/variants/librtrb_w32.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_api_contract.py -m gpu -x -q 2>&1 | tail -4
