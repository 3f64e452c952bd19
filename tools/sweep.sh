#!/bin/bash
# usage: tools/sweep.sh "<nvcc -D flags>" ...   (rebuilds the fast TU with each flag set on the GPU box, prints config 2/3/5 times)
for flags in "$@"; do
  (cd raytracing_rb_b200/csrc && touch rtrb_trace_fast.cu && make -s NVCC="nvcc $flags" > /dev/null 2>&1)
  echo "== $flags"
  python tools/perf_configs.py fast64 2>&1 | grep -E "config (2|3|5)"
done
