#!/bin/bash
# ncu --set full capture of the trace kernel of each config; summaries and the per-instruction source page come back,
# the (large) .ncu-rep files stay on the box.  Usage: r2_ncu.sh <tag> <config ids...>
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
for c in "$@"; do
  case $c in
    2) ARGS="2 --frames 2";;
    3) ARGS="3 --frames 2";;
    4) ARGS="4 --frames 2";;
    5) ARGS="5 --width 480 --height 270 --spp 64 --frames 2";;
    5s) ARGS="5 --width 480 --height 270 --spp 4 --frames 2";;
  esac
  python tools/run_config.py $ARGS > $O/plain_c$c.log 2>&1 &&
  $NCU -k regex:trace_pre_ -s 1 -c 1 -o $O/c$c -f python tools/run_config.py $ARGS > $O/ncu_c$c.log 2>&1
  python tools/ncu_summary.py $O/c$c.ncu-rep "config $c ($ARGS) $TAG" > $O/summary_c$c.txt 2>&1
  ncu -i $O/c$c.ncu-rep --page source --csv 2>/dev/null | gzip > $O/source_c$c.csv.gz
  ncu -i $O/c$c.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/raw_c$c.csv.gz
  rm -f $O/c$c.ncu-rep
  cat $O/plain_c$c.log
done
ls -la $O
