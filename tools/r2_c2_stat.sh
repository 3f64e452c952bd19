#!/bin/bash
# robust config-2 kernel timing: median / min of the trace kernel over 40 frames, per library build
for lib in "$@"; do
  if [ "$lib" = "main" ]; then unset RTRB_B200_LIB; else export RTRB_B200_LIB=$PWD/raytracing_rb_b200/csrc/variants/$lib; fi
  python tools/run_config.py 2 --frames 45 | python -c "
import sys,re,statistics
v=[float(re.search(r'trace ([0-9.]+) ms',l).group(1)) for l in sys.stdin if 'trace' in l][5:]
print('$lib', 'median %.4f min %.4f max %.4f n %d' % (statistics.median(v), min(v), max(v), len(v)))"
done
