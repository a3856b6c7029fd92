#!/bin/bash
# Final-tree check: full GPU test suite, smoke, the default bench line, per-kernel event times
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
timeout 400 python bench.py 2>/dev/null | grep "^{" > $O/bench.json
for i in 1 2; do timeout 120 python profiles/tools/kernel_times.py 1024 60 | tail -1; done > $O/kernel_times.txt
timeout 200 python profiles/tools/kernel_times.py 8192 30 16000 1 3.5 | tail -1 >> $O/kernel_times.txt
cat $O/pytest.txt $O/smoke.txt $O/kernel_times.txt; cut -c1-300 $O/bench.json
