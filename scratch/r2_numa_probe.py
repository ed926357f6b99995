"""Host-link placement probe (torchrun, one rank per GPU): per-rank concurrent H2D / D2H copy rates of page-locked
buffers allocated (a) wherever the process happens to run, (b) after binding the process to the CPUs of the NUMA node
its GPU hangs off.  Prints the topology it found."""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def gpu_numa_node(index):
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        return bus, node
    except Exception as e:  # noqa: BLE001
        return repr(e), -1


def node_cpus(node):
    try:
        spec = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:  # noqa: BLE001
        return set()


bus, node = gpu_numa_node(lr)
allowed = os.sched_getaffinity(0)
if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout, flush=True)
        print("nodes:", sorted(os.listdir("/sys/devices/system/node"))[:12], flush=True)
    except Exception as e:  # noqa: BLE001
        print("topo:", e)
nb = 1 << 29
d = torch.empty(nb // 8, dtype=torch.int64, device=dev)


def rates(tag):
    h = torch.empty(nb // 8, dtype=torch.int64, pin_memory=True)
    h.zero_()
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        out[name] = round(4 * nb / (time.perf_counter() - t0) / 1e9, 1)
    if world > 1:
        dist.barrier()
    print(f"rank {rank} gpu {bus} numa {node} allowed_cpus {len(allowed)} {tag}: {out}", flush=True)
    del h


rates("default placement")
cpus = node_cpus(node) & allowed if node >= 0 else set()
if cpus:
    os.sched_setaffinity(0, cpus)
rates(f"bound to node {node} ({len(cpus)} cpus)" if cpus else "no binding possible")
if world > 1:
    dist.destroy_process_group()
