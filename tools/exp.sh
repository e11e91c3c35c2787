timeout 300 python tools/train_probe.py 2>&1 | tail -8
