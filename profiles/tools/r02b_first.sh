#!/bin/bash
# Re-entry capture: GPU tests, per-kernel times, K4 phase table (developer build), ncu source-level K4
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python profiles/tools/kernel_times.py 1024 60 > $O/kt_1024.txt 2>&1
SPEEDY_K4_CHAIN=1 python profiles/tools/kernel_times.py 1024 60 > $O/kt_1024_chain.txt 2>&1
python profiles/tools/kernel_times.py 8192 30 16000 1 3.5 > $O/kt_8192.txt 2>&1
cat $O/kt_*.txt
SPEEDY_B200_LIB=$PWD/scratch/timing/libspeedy_b200.so python profiles/tools/k4_phases.py 1024 60 2.0 $O/k4_phases_onewarp.json > /dev/null 2> $O/k4_phases.err
python profiles/tools/show_phases.py $O/k4_phases_onewarp.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_a.json 2> $O/bench_a.err; tail -c 1200 $O/bench_a.json
bash profiles/tools/ncu_k4.sh
echo done
