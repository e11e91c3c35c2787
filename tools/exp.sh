timeout 900 python -m pytest tests -x -q -m gpu -k "pipelined or graph or host or forward_empty" > gpurun_out/t35.log 2>&1; tail -3 gpurun_out/t35.log
timeout 300 python bench.py --no-cpu > gpurun_out/exp21.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"e2e": {[^}]*}' gpurun_out/exp21.log
