timeout 900 python -m pytest tests -x -q -m gpu -k "anchor or forward or headline or large or training or grad" > gpurun_out/t34.log 2>&1; tail -3 gpurun_out/t34.log
for i in 1 2; do timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp20_$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*\|"stage_ms": {[^}]*}' gpurun_out/exp20*.log
