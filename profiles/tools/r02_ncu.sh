#!/bin/bash
# ncu --set full capture (source-level) of K1, K2, K4 (one launch each covers the whole 1024 x 60 s step)
# and the K4 per-phase cycle tables from the -DK4_TIMING developer build.
mkdir -p gpurun_out; O=gpurun_out
if [ -f scratch/timing/libspeedy_b200.so ]; then
  SPEEDY_B200_LIB=$PWD/scratch/timing/libspeedy_b200.so python profiles/tools/k4_phases.py 1024 60 2.0 $O/k4_phases_onewarp.json > /dev/null 2> $O/k4_phases.err
  SPEEDY_K4_PIPELINE=1 SPEEDY_B200_LIB=$PWD/scratch/timing/libspeedy_b200.so python profiles/tools/k4_phases.py 1024 60 2.0 $O/k4_phases_pipeline.json > /dev/null 2>> $O/k4_phases.err
fi
SPEEDY_B200_WRITE_PARTS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k4_sonic|k1_spectral|k2_tension' --launch-skip 12 --launch-count 4 -o $O/prof_r2 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/prof_r2.log 2>&1
ncu -i $O/prof_r2.ncu-rep --page source --csv --print-source sass --kernel-name regex:k4_sonic --launch-count 1 > $O/prof_r2_k4_sass.csv 2>/dev/null
echo done
