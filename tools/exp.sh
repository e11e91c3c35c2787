timeout 600 python -m pytest tests -x -q -m gpu -k "project or forward or headline or large" > gpurun_out/t31.log 2>&1; tail -5 gpurun_out/t31.log
for i in 1 2; do timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp17_$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*\|"stage_ms": {[^}]*}' gpurun_out/exp17*.log
