python -m pytest tests -x -q -m gpu -k "graph" > gpurun_out/t26.log 2>&1; tail -5 gpurun_out/t26.log
for i in 1 2 3 4 5 6 7 8; do python bench.py --no-cpu --no-e2e > gpurun_out/exp11_a$i.log 2>&1; done
for i in 1 2 3; do python bench.py --no-cpu --no-e2e --no-graph > gpurun_out/exp11_b$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/exp11*.log
