timeout 600 python -m pytest tests -x -q -m gpu -k "decode or sharding" > gpurun_out/t46.log 2>&1; tail -2 gpurun_out/t46.log
