#!/bin/bash
# ncu --set full source-level capture of the K4 kernel of one 1024 x 60 s write (one launch)
mkdir -p gpurun_out; O=gpurun_out
SPEEDY_B200_WRITE_PARTS=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'k4_' --launch-skip 2 --launch-count 1 -o $O/prof_k4 -f \
  python profiles/tools/kernel_times.py 1024 60 > $O/prof_k4.log 2>&1
ncu -i $O/prof_k4.ncu-rep --page source --csv --print-source sass > $O/prof_k4_sass.csv 2>/dev/null
ncu -i $O/prof_k4.ncu-rep --page raw --csv > $O/prof_k4_raw.csv 2>/dev/null
tail -2 $O/prof_k4.log
