timeout 600 python -m pytest tests -x -q -m gpu -k "aff or forward_matches or large" > gpurun_out/t48.log 2>&1; tail -2 gpurun_out/t48.log
timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp36.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"col_softmax": [0-9.]*' gpurun_out/exp36.log
