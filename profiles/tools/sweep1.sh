python profiles/tools/step_time.py 2>&1 | tail -1
SPEEDY_K4_CHAIN_N=2048 python profiles/tools/step_time.py 2>&1 | tail -1
SPEEDY_K4_CHAIN=0 python profiles/tools/step_time.py 2>&1 | tail -1
SPEEDY_K4_CHAIN_N=2048 python profiles/tools/kernel_times.py 1024 60
SPEEDY_K4_CHAIN_MAX=100000 python profiles/tools/kernel_times.py 8192 30 16000 1 3.5
SPEEDY_K4_CHAIN_MAX=100000 python profiles/tools/step_time.py 8192 30 2>&1 | tail -1
python profiles/tools/step_time.py 8192 30 2>&1 | tail -1
