#!/bin/bash
# Round 2, step 0: timings and ncu --set full captures of the ray-tree kernels as shipped at the end of round 1.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
python tools/run_config.py 1 --frames 4 > $O/times.txt 2>&1
python tools/run_config.py 2 --frames 4 >> $O/times.txt 2>&1
python tools/run_config.py 3 --frames 4 >> $O/times.txt 2>&1
python tools/run_config.py 4 --frames 4 >> $O/times.txt 2>&1
python tools/run_config.py 5 --width 480 --height 270 --spp 64 --frames 3 >> $O/times.txt 2>&1
python tools/run_config.py 5 --frames 2 >> $O/times.txt 2>&1
cat $O/times.txt
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:trace_pre_fast -s 1 -c 1 -o $O/c3 -f python tools/run_config.py 3 --frames 2 > $O/ncu_c3.log 2>&1
$NCU -k regex:trace_pre_fast -s 1 -c 1 -o $O/c4 -f python tools/run_config.py 4 --frames 2 > $O/ncu_c4.log 2>&1
$NCU -k regex:trace_pre_fast -s 1 -c 1 -o $O/c5 -f python tools/run_config.py 5 --width 480 --height 270 --spp 64 --frames 2 > $O/ncu_c5.log 2>&1
$NCU -k regex:trace_pre_fast -s 1 -c 1 -o $O/c2 -f python tools/run_config.py 2 --frames 2 > $O/ncu_c2.log 2>&1
ls -la $O
