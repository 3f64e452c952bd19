#!/bin/bash
# config 5 only (many-sphere scene), for BVH experiments.  Usage: r2_c5.sh <tag> <lib|main> ...
TAG=$1; shift; O=gpurun_out/$TAG; mkdir -p $O
for lib in "$@"; do
  if [ "$lib" = "main" ]; then unset RTRB_B200_LIB; else export RTRB_B200_LIB=$PWD/raytracing_rb_b200/csrc/variants/$lib; fi
  echo "=== $lib" | tee -a $O/times.txt
  for a in "5 --width 480 --height 270 --spp 4 --frames 3" "5 --width 480 --height 270 --spp 64 --frames 3" "5 --frames 2"; do
    timeout 120 python tools/run_config.py $a 2>&1 | tail -1 | tee -a $O/times.txt
  done
done
unset RTRB_B200_LIB
