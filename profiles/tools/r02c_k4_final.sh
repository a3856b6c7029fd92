#!/bin/bash
# Final-tree evidence: full GPU test suite, smoke, the bench line, and one ncu --set full source-level
# capture of K4 (one 1024 x 60 s write in one launch); the report stays in /tmp, only CSV exports travel.
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
timeout 400 python bench.py 2>/dev/null | grep "^{" > $O/bench.json
SPEEDY_B200_WRITE_PARTS=1 timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:'k4_sonic' --launch-skip 2 --launch-count 1 -o /tmp/prof_k4 -f \
  python profiles/tools/kernel_times.py 1024 60 > $O/prof_k4.log 2>&1
ncu -i /tmp/prof_k4.ncu-rep --page source --csv --print-source sass > $O/prof_k4_sass.csv 2>/dev/null
ncu -i /tmp/prof_k4.ncu-rep --page raw --csv > $O/prof_k4_raw.csv 2>/dev/null
ls -la $O; cat $O/pytest.txt $O/smoke.txt; cut -c1-400 $O/bench.json; tail -3 $O/prof_k4.log
