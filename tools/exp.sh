timeout 900 python -m pytest tests -x -q -m gpu -k "side_stream or forward or graph or pipelined or training or headline" > gpurun_out/t45.log 2>&1; tail -2 gpurun_out/t45.log
for i in 1 2 3 4; do timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/exp31_$i.log 2>&1; done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/exp31*.log
