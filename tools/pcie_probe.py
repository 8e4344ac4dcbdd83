"""H2D bandwidth from pinned host memory for several copy sizes and CPU affinities (diagnostic).
Answers: is bench.py's e2e leg (8.5 MB per step) PCIe-bound, and does NUMA placement matter?"""
import os
import sys
import time

import torch


def bw(nbytes, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    os.system("nvidia-smi topo -m 2>&1 | head -20")
    os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv")
    os.system("lscpu | grep -i -E 'numa|model name|socket' | head")
    all_cpus = sorted(os.sched_getaffinity(0))
    halves = {"all": all_cpus, "first_half": all_cpus[:len(all_cpus) // 2], "second_half": all_cpus[len(all_cpus) // 2:]}
    for name, cpus in halves.items():
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        row = [f"{bw(n):.1f}" for n in (1 << 15, 1 << 20, 4 << 20, 8 << 20, 64 << 20)]
        print(f"affinity={name:11s} H2D GB/s at 32K/1M/4M/8M/64M:", " ".join(row), flush=True)
    os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    sys.exit(main())
