timeout 600 python -m pytest tests/test_tracker.py -x -q > gpurun_out/t38.log 2>&1; tail -8 gpurun_out/t38.log
