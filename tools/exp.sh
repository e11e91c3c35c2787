python tools/bench_shared_conv.py --maps 8 --hw 180 > gpurun_out/conv1.log 2>&1; tail -1 gpurun_out/conv1.log
python tools/bench_shared_conv.py --maps 4 --hw 512 --iters 5 > gpurun_out/conv2.log 2>&1; tail -1 gpurun_out/conv2.log
