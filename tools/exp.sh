timeout 600 python -m pytest tests -x -q -m gpu -k "gather or pipelined or all_zero" > gpurun_out/t43.log 2>&1; tail -2 gpurun_out/t43.log
timeout 300 python bench.py --no-cpu > gpurun_out/exp28.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"e2e": {[^}]*}' gpurun_out/exp28.log | cut -c1-200
timeout 300 python tools/pcie_probe.py 2>&1 | tail -4
