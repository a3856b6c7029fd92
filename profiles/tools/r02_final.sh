#!/bin/bash
# Round-2 final capture, one gpurun call: GPU tests, bench lines of configs 2..5 and the reference arm,
# per-kernel times, launch list, ncu --set full of the three kernels of a step, the 48 kHz stereo and
# streaming-step captures (exported to CSV on the box: the reports themselves are too large to bring back).
#   gpurun --timeout 2400 -- 'bash profiles/tools/r02_final.sh'
mkdir -p gpurun_out; O=gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
  tail -3 $O/pytest_gpu.log
fi
python bench.py --steps 10 --warmup 3 > $O/bench_r2.json 2> $O/bench_r2.err; tail -c 300 $O/bench_r2.json
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_c$c.json 2> $O/bench_r2_c$c.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_ref.json 2> $O/bench_r2_ref.err
python profiles/tools/kernel_times.py 1024 60 > $O/kt.txt 2>&1
SPEEDY_K1_TC=0 python profiles/tools/kernel_times.py 1024 60 >> $O/kt.txt 2>&1
SPEEDY_K4_CHAIN=1 python profiles/tools/kernel_times.py 1024 60 >> $O/kt.txt 2>&1
python profiles/tools/kernel_times.py 8192 30 16000 1 3.5 >> $O/kt.txt 2>&1
python profiles/tools/kernel_times.py 1024 10 48000 2 1.5 >> $O/kt.txt 2>&1
cat $O/kt.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_r2.log 2>&1
T=/tmp/ncu_r2; mkdir -p $T
SPEEDY_B200_WRITE_PARTS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k4_sonic|k1_dft16|k2_tension' --launch-skip 6 --launch-count 4 -o $T/prof_r2 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/prof_r2.log 2>&1
ncu -i $T/prof_r2.ncu-rep --page raw --csv > $O/prof_r2_raw.csv 2>/dev/null
SPEEDY_B200_WRITE_PARTS=1 timeout 600 ncu --set full --clock-control none \
  -k regex:'k4_sonic|k1_spectral|k2_tension' --launch-skip 3 --launch-count 3 -o $T/prof_r2_48k -f \
  python profiles/tools/kernel_times.py 1024 10 48000 2 1.5 > $O/prof_r2_48k.log 2>&1
ncu -i $T/prof_r2_48k.ncu-rep --page raw --csv > $O/prof_r2_48k_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none \
  -k regex:'k4_sonic|k1_dft16|k2_tension|tail_kernel|read_' --launch-skip 40 --launch-count 6 -o $T/prof_r2_stream -f \
  python profiles/tools/stream_step.py > $O/prof_r2_stream.log 2>&1
ncu -i $T/prof_r2_stream.ncu-rep --page raw --csv > $O/prof_r2_stream_raw.csv 2>/dev/null
du -sh $O; echo done
