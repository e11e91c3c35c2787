timeout 300 python -m pytest tests -x -q -m gpu -k "aff" -s > gpurun_out/t29.log 2>&1; tail -8 gpurun_out/t29.log
for i in 1; do python bench.py --no-cpu --no-e2e > gpurun_out/exp14_$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*\|"stage_ms": {[^}]*}' gpurun_out/exp14*.log
