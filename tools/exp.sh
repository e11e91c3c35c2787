timeout 600 python -m pytest tests -x -q -m gpu -k "pairwise or forward_matches or headline" > gpurun_out/t39.log 2>&1; tail -3 gpurun_out/t39.log
for i in 1 2; do timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp24_$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*\|"pairwise": [0-9.]*' gpurun_out/exp24*.log
