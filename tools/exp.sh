timeout 600 python -m pytest tests -x -q -m gpu -k "pairwise or forward or headline or large" > gpurun_out/t30.log 2>&1; tail -5 gpurun_out/t30.log
for i in 1 2; do timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp15_$i.log 2>&1; done
timeout 120 python bench.py --no-cpu --no-e2e --flags 0x20 > gpurun_out/exp15_bf16.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"stage_ms": {[^}]*}' gpurun_out/exp15*.log
