#!/bin/bash
# One gpurun call: GPU tests, the bench lines of configs 2..5, the reference arm, the launch
# list and the K4 per-phase cycle tables.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash profiles/tools/r02_capture.sh'
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_r2.json 2> $O/bench_r2.err; tail -c 1500 $O/bench_r2.json
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_c$c.json 2> $O/bench_r2_c$c.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_ref.json 2> $O/bench_r2_ref.err
python profiles/tools/kernel_times.py 1024 60 > $O/kt_1024.txt 2>&1
SPEEDY_K4_PIPELINE=1 python profiles/tools/kernel_times.py 1024 60 > $O/kt_1024_pipe.txt 2>&1
python profiles/tools/kernel_times.py 8192 30 16000 1 3.5 > $O/kt_8192.txt 2>&1
cat $O/kt_*.txt
if [ -f scratch/timing/libspeedy_b200.so ]; then
  SPEEDY_B200_LIB=$PWD/scratch/timing/libspeedy_b200.so python profiles/tools/k4_phases.py 1024 60 2.0 $O/k4_phases_onewarp.json > /dev/null 2> $O/k4_phases.err
  SPEEDY_K4_PIPELINE=1 SPEEDY_B200_LIB=$PWD/scratch/timing/libspeedy_b200.so python profiles/tools/k4_phases.py 1024 60 2.0 $O/k4_phases_pipeline.json > /dev/null 2>> $O/k4_phases.err
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_r2.log 2>&1
echo done
