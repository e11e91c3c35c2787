"""Is the pinned host memory NUMA-local to the GPU? Prints the topology and the H2D / zero-copy bandwidth with the
process bound to each NUMA node's CPUs while the pinned buffer is allocated and first touched."""
import glob
import os
import subprocess
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def main():
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
    allowed = sorted(os.sched_getaffinity(0))
    print("allowed cpus:", len(allowed), allowed[:4], "...", allowed[-4:])
    print("numa nodes:", {k: (len(v), v[:2], v[-2:]) for k, v in nodes.items()})
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(0)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        print("nvml ideal cpus:", len(ideal), ideal[:4], "...", ideal[-4:])
    except Exception as e:  # noqa: BLE001
        print("nvml affinity unavailable:", e)
    torch.cuda.init()
    dev = torch.device("cuda:0")
    n = 256 << 20
    dst = torch.empty(n // 4, dtype=torch.float32, device=dev)
    for node, cpus in list(nodes.items()) + [(-1, allowed)]:
        use = sorted(set(cpus) & set(allowed))
        if not use:
            print("node", node, "no allowed cpus")
            continue
        os.sched_setaffinity(0, use)
        src = torch.empty(n // 4, dtype=torch.float32).pin_memory()
        src.fill_(1.0)
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        print("node %d: H2D %.1f GB/s" % (node, 8 * n / e0.elapsed_time(e1) / 1e6))
        del src
    os.sched_setaffinity(0, allowed)


if __name__ == "__main__":
    main()
