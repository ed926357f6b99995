"""Sharded four-step NTT: parity against the single-GPU transform + timing.  Run plain (world 1)
or under torchrun."""
import os, sys, json, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, torch, torch.distributed as dist
import b200zk
from b200zk import sharding
from b200zk.api import _ptr, fr_limbs, FR_ROOT_OF_UNITY, FR_MODULUS

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
b200zk.init(lr); lib = b200zk.load()
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["nccl"]
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [12, 20, 24]
for mode in modes:
    for k in ks:
        n = 1 << k
        omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
        full = torch.empty(n * 4, dtype=torch.int64, device=dev)
        b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(full.data_ptr()), n, 0xA11CE000 + k, 0))
        ref = full.clone()
        b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(ref.data_ptr()), n, 1, k, _ptr(fr_limbs(omega)), None, None))
        torch.cuda.synchronize()
        log_n1 = sharding.four_step_split(k, world)
        n1, n2 = 1 << log_n1, 1 << (k - log_n1); m = n2 // world
        # column block on the device: [n1][n2] -> [:, r m:(r+1) m]
        x0 = full.view(n1, n2, 4)[:, rank * m:(rank + 1) * m].contiguous().view(-1)
        ops = sharding.DeviceFourStep(k, world, rank, dev, mode=mode)
        x = x0.clone()
        rows = sharding.sharded_best_fft(x, k, omega, ops, world, rank)
        torch.cuda.synchronize()
        # expected rows: A[i1 + n1 i2] for i1 in my range
        r0 = rank * (n1 // world)
        want = ref.view(n2, n1, 4)[:, r0:r0 + n1 // world].permute(1, 0, 2).contiguous().view(-1)
        ok = bool(torch.equal(rows.view(-1), want))
        # timing
        reps = 5
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        xs = [x0.clone() for _ in range(reps)]
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0.record()
        for i in range(reps):
            sharding.sharded_best_fft(xs[i], k, omega, ops, world, rank)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        okt = torch.tensor([1 if ok else 0], device=dev)
        if world > 1: dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"k": k, "world": world, "mode_requested": mode, "mode": ops.mode, "parity": bool(okt.item()),
                              "ms": round(float(ms.item()), 4), "melem_per_s": round(n / float(ms.item()) / 1e3, 1),
                              "p2p_error": getattr(ops, "p2p_error", None)}), flush=True)
        del ops, x, xs, full, ref
if world > 1:
    dist.destroy_process_group()
