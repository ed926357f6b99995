#!/usr/bin/env python
"""bench.py — the BN254 MSM + Fr NTT hot path of the Halo2/KZG prover on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--k 24] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): "standalone BN254 MSM and Fr NTT sweep k=16..26 on
synthetic scalars/points".  One *step* = one KZG commitment `ParamsKZG::commit_lagrange`
(= best_multiexp of 2^k seeded scalars against the 2^k-point SRS, which is registered on
the device once, as it is fixed for the life of the params) plus one `best_fft` over 2^k
seeded scalars, per GPU.  The same MSM through plain `best_multiexp` (bases passed per
call, no per-window table) is reported in `msm_unregistered`.  With N GPUs the MSM
is the north-star point-range split: rank r owns points [r*2^k, (r+1)*2^k) of one
N*2^k-point MSM, partial sums are combined by a one-point-per-rank gather (NCCL
all_gather of 96 B) + a fold on rank 0; the NTT columns are independent per rank.
Per-GPU work is fixed, so scaling is "weak".

Printed line: see the contract in the task statement; `value` is Mpts/s through the
whole step with inputs resident in HBM; `e2e` is the same through the host-buffer C-ABI
calls (`b200zk_msm_g1_registered`, `b200zk_ntt`) with pinned host buffers, H2D/D2H inside
the timed region; `roofline` is the dominant kernel (bucket accumulation) against the measured
integer-pipe modmul peak, `roofline_ntt` the NTT against the measured HBM copy peak.
`--impl reference` times the restated CPU baseline (oracle/halo2_oracle.c; the
reference's Rust prover cannot be built here: no cargo, dependencies not vendored).
`--sweep` prints the k = 16..26 table used in DESIGN.md (not a driver line).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))

import numpy as np  # noqa: E402

METRIC = "MSM Mpts/s (step = BN254 KZG commit_lagrange MSM 2^k + Fr best_fft 2^k)"
UNIT = "Mpts/s"
SEED_S, SEED_P = 0xA11CE000, 0xBA5E0000


def fr_limbs(x: int) -> np.ndarray:
    from b200zk.api import fr_limbs as f
    return f(x)


def omega_for(k: int) -> int:
    from b200zk.api import FR_MODULUS, FR_ROOT_OF_UNITY, FR_S
    w = FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR_MODULUS
    return w


def msm_modmul_model(npairs: int) -> float:
    """Algorithmic work of the bucket-accumulation kernel: one XYZZ mixed addition
    (8M + 2S = 10 field multiplications) per (point, window) pair."""
    return 10.0 * npairs


def ntt_alg_bytes(k: int) -> float:
    """SURVEY.md section 8 d: 64 * n * ceil(k / 12)."""
    return 64.0 * (1 << k) * max(1, -(-k // 12))


# ---- exact check of a commitment against the known discrete logs of the synthetic SRS ----------
# The synthetic bases are P_i = [t_i] G with t_i = splitmix64(seed << 32 + i) | 1 (the stream
# b200zk_gen_points_dev documents), so sum_i s_i P_i = [sum_i s_i t_i mod r] G.  Plain Python / numpy
# integers: independent of the library and of oracle/.
_FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _splitmix64_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def point_discrete_logs(seed: int, start: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64((seed << 32) & ((1 << 64) - 1)) + np.uint64(start)
        return _splitmix64_np(base + np.arange(n, dtype=np.uint64)) | np.uint64(1)


def dot_mod_r(s_limbs: np.ndarray, t: np.ndarray) -> int:
    """sum_i s_i t_i mod r for s (n, 4) u64 limbs and t (n,) u64, exactly: 32-bit pieces of s against
    16-bit pieces of t, 65536 rows at a time (48-bit products: a block sum stays below 2^64)."""
    n = s_limbs.shape[0]
    s32 = np.ascontiguousarray(s_limbs).view(np.uint32).reshape(n, 8).astype(np.uint64)
    t16 = np.ascontiguousarray(t).view(np.uint16).reshape(n, 4).astype(np.uint64)
    total = 0
    for lo in range(0, n, 1 << 16):
        m = s32[lo:lo + (1 << 16)].T @ t16[lo:lo + (1 << 16)]
        for a in range(8):
            for b in range(4):
                total += int(m[a, b]) << (32 * a + 16 * b)
    return total % _FR


def _g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % _FQ == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, _FQ) % _FQ
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, _FQ) % _FQ
    x3 = (lam * lam - x1 - x2) % _FQ
    return x3, (lam * (x1 - x3) - y1) % _FQ


def generator_mul(scalar: int):
    """[scalar] (1, 2) on y^2 = x^3 + 3 over Fq, affine (None = identity)."""
    acc, add = None, (1, 2)
    scalar %= _FR
    while scalar:
        if scalar & 1:
            acc = _g1_add(acc, add)
        add = _g1_add(add, add)
        scalar >>= 1
    return acc


def jacobian_limbs_to_affine(limbs12: np.ndarray):
    """12 u64 limbs (Montgomery Jacobian G1, the MSM's return type) -> canonical affine (x, y) or None."""
    rinv = pow(1 << 256, -1, _FQ)
    v = [int.from_bytes(np.ascontiguousarray(limbs12[4 * i: 4 * i + 4]).tobytes(), "little") * rinv % _FQ for i in range(3)]
    x, y, z = v
    if z == 0:
        return None
    zi = pow(z, -1, _FQ)
    return x * zi * zi % _FQ, y * zi * zi * zi % _FQ


def commitment_matches_discrete_logs(point12: np.ndarray, dot: int) -> bool:
    """dot = sum s_i t_i over the *Montgomery representatives* s_i (what the limbs hold): the scalars the
    MSM uses are s_i / R."""
    return jacobian_limbs_to_affine(point12) == generator_mul(dot * pow(1 << 256, -1, _FR) % _FR)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (t, sm_mhz, max_mhz, watts, reasons bitmask) from NVML
        self._stop = threading.Event()

    # NVML clocks-event-reason bits (nvml.h): sw power cap 0x4, hw slowdown 0x8, sw thermal 0x20, hw thermal 0x40
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), float(sm), float(mx), pw, int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.nvml = pynvml
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin: float = 0.0, t_end: float = 1e300) -> dict:
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            inside = [x for x in self.samples if t_begin <= x[0] <= t_end]
            window = "timed region"
            if not inside:
                inside, window = list(self.samples), "warm-up + timed region"
            reasons = sorted(nm for nm, bit in self.REASON_BITS.items() if any(x[4] & bit for x in inside))
            return {"sm_mhz": statistics.median(x[1] for x in inside) if inside else None,
                    "sm_max_mhz": max((x[2] for x in inside), default=None),
                    "power_w_max": max((x[3] for x in inside), default=None), "samples": len(inside),
                    "window": window, "source": "NVML, 5 ms period", "reasons": reasons}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (ts, ln) in self.lines if t_begin <= ts <= t_end + 0.15]
        window = "timed region"
        if not inside:  # region shorter than one sampling period: fall back to warm-up + timed region
            inside = [ln for (_, ln) in self.lines]
            window = "warm-up + timed region (timed region shorter than one nvidia-smi period)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window,
                "source": "nvidia-smi -lms 100", "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_step(co, k: int, threads: int, scalars=None, points=None):
    """One step of the restated CPU baseline at size 2^k; returns (seconds, msm_s, fft_s)."""
    n = 1 << k
    if scalars is None:
        scalars = co.gen_scalars(SEED_S + k, n)
        points = co.gen_points(SEED_P + k, n, threads=threads)
    w = fr_limbs(omega_for(k))
    t0 = time.perf_counter()
    cpu_step.last_point = co.best_multiexp(scalars, points, threads)
    t1 = time.perf_counter()
    cpu_step.last_fft = co.best_fft(scalars, w, k, threads)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1


def run_reference(args) -> None:
    """--impl reference: the restated CPU baseline on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle as co
    co.build()
    threads = co.host_threads()
    # the size this arm prints is the size it runs: the GPU arm's own 2^k (about 10 s per step at k = 24 on
    # 16 cores); --cpu-k runs, and reports, a smaller one
    k = args.cpu_k or args.k
    args.k = k
    n = 1 << k
    scalars = co.gen_scalars(SEED_S + k, n)
    points = co.gen_points(SEED_P + k, n, threads=threads)
    for _ in range(args.warmup):
        cpu_step(co, k, threads, scalars, points)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(co, k, threads, scalars, points)
    dt = time.perf_counter() - t0
    value = args.steps * n / dt / 1e6
    sample = f"each step = one whole best_multiexp + best_fft at 2^{k}, the GPU arm's seeds and size"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (Montgomery, 254-bit)",
        "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated CPU baseline (C port of halo2_proofs best_multiexp/best_fft); the reference's Rust "
                "prover cannot be built in this image (no cargo, git dependencies not vendored)",
    }
    print(json.dumps(line), flush=True)


def ntt_modmul_per_element(k: int) -> float:
    """Field multiplications per element that csrc/ntt.cuh actually executes for one 2^k transform (counted from
    the kernel, not k/2 + twiddles): per pass of 2^b rows the rounds are radix 8, 8, ..., then 2^(b mod 3); a
    radix-8 round multiplies 5 of its 12 butterfly outputs (the others meet the trivial root), a radix-4 round 1 of
    4; between rounds every element whose digit c and lower index lo are both non-zero takes one in-tile twiddle;
    a pass boundary costs one multiplication per element with a direct table (boundaries of <= 2^20 twiddles) and
    two with the lo / hi tables.  12.0 at k = 24 — what radix-2 needs (k/2), not more."""
    npass = max(1, -(-k // 9))
    bits = [k // npass + (1 if p < k % npass else 0) for p in range(npass)]
    per, log_i = 0.0, 0
    for p, b in enumerate(bits):
        rounds = [3] * (b // 3) + ([b % 3] if b % 3 else [])
        for q, r in enumerate(rounds):
            per += {3: 5 / 8, 2: 1 / 4, 1: 0.0}[r]
            later = sum(rounds[q + 1:])
            if q + 1 < len(rounds):
                per += (1 - 2.0 ** -r) * (1 - 2.0 ** -later)
        if p + 1 < npass:
            nonzero = (1 - 2.0 ** -b) * (1 - 2.0 ** -(k - log_i - b))     # exponent i_p * J is zero when either is
            per += (1 if k - log_i <= 20 else 2) * nonzero
        log_i += b
    return per


def workload_config(args, world: int) -> dict:
    return {
        "workload": f"BASELINE.json configs[1]: standalone BN254 MSM + Fr NTT at k={args.k} "
                    f"(2^{args.k} points and scalars per GPU, seeds 0xA11CE000+k / 0xBA5E0000+k)",
        "k": args.k, "points_per_gpu": 1 << args.k, "global_points": world << args.k,
        "parallelism": f"point-range split x{world}, one-point-per-rank gather" if world > 1 else "single GPU",
        "l2": "inputs (96 B/point + 32 B/NTT element, >= 2 GiB at k=24) exceed the 126 MB L2; no flush needed",
        "streams": ("transform on a second, lower-priority stream under the commit (its kernel time below is the "
                    "stretched, concurrent one)") if getattr(args, "overlap", False)
                   else "commit and transform back to back on one stream",
    }


# ----------------------------------------------------------------------------- GPU arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--k", type=int, default=24, help="log2 points per GPU")
    ap.add_argument("--cpu-k", type=int, default=0,
                    help="log2 size of the CPU legs (0: min(k, 24) for cpu_baseline on the step's own inputs, k for "
                         "--impl reference; the size run is the size reported)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", action="store_true",
                    help="issue the transform on a second, lower-priority stream under the commit")
    ap.add_argument("--sweep", action="store_true", help="print the k=16..26 table instead of the driver line")
    ap.add_argument("--no-proof-shape", action="store_true",
                    help="skip the RSA-SHA256-shaped create_proof hot-path pass (BASELINE.json configs[2] stand-in)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import b200zk
    from b200zk.api import _ptr

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200zk product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    b200zk.init(local_rank)
    lib = b200zk.load()
    dev = torch.device("cuda", local_rank)

    if args.sweep:
        sweep(args, torch, b200zk, lib, dev)
        return

    k, n = args.k, 1 << args.k
    stream = torch.cuda.Stream(device=dev, priority=-1)
    st = C.c_void_p(stream.cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())

    # ---- synthetic inputs, generated in HBM (rank r owns the r-th slice of the global MSM)
    d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
    d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
    d_ntt = torch.empty(n * 4, dtype=torch.int64, device=dev)
    b200zk.check(lib.b200zk_gen_scalars_dev(vp(d_scal), n, SEED_S + k, rank * n))
    b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), n, SEED_P + k, rank * n))
    d_ntt.copy_(d_scal)
    d_pt = torch.zeros(12, dtype=torch.int64, device=dev)
    gathered = torch.zeros(world * 12, dtype=torch.int64, device=dev)
    # ParamsKZG: the SRS slice of this rank is uploaded / tabulated once, outside the timed region
    t_reg = time.perf_counter()
    h_bases_np = d_base.cpu().numpy().view(np.uint64).reshape(n, 8)
    handle = C.c_uint64(0)
    b200zk.check(lib.b200zk_bases_register(_ptr(h_bases_np), n, C.byref(handle)))
    torch.cuda.synchronize()
    t_reg = time.perf_counter() - t_reg
    omega = fr_limbs(omega_for(k))
    peak_modmul = b200zk.modmul_peak(4096)
    result_holder = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ntt_ev = []
    # The commit and the transform of a step are independent (as the column commitments and transforms of
    # create_proof are).  With --overlap the transform is issued on a second, lower-priority stream right
    # after the commit's kernels are queued: its CTAs fill the SM slots the commit leaves idle (the
    # latency-bound bucket reduction runs at ~12 % warp occupancy), and the step ends when both are done.
    lo_stream = torch.cuda.Stream(device=dev, priority=0)
    st_lo = C.c_void_p(lo_stream.cuda_stream)

    def step(timed: bool):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            if args.overlap:
                fork = torch.cuda.Event()
                fork.record(stream)
            b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, 1, n, vp(d_pt), st))
            if args.overlap:
                lo_stream.wait_event(fork)                 # not before this step began
                e0.record(lo_stream)
                b200zk.check(lib.b200zk_ntt_dev(vp(d_ntt), n, 1, k, _ptr(omega), None, st_lo))
                e1.record(lo_stream)
            if world > 1:
                dist.all_gather_into_tensor(gathered, d_pt)
                if rank == 0:
                    pts = gathered.cpu().numpy().view(np.uint64).reshape(world, 12)
                    result_holder["point"] = b200zk.g1_sum(np.ascontiguousarray(pts))
            if args.overlap:
                stream.wait_event(e1)                      # join
            else:
                e0.record(stream)
                b200zk.check(lib.b200zk_ntt_dev(vp(d_ntt), n, 1, k, _ptr(omega), None, st))
                e1.record(stream)
            if timed:
                ntt_ev.append((e0, e1))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(False)
    barrier()
    b200zk.check(lib.b200zk_msm_profile(1))
    launches0 = b200zk.kernel_launches()
    stage_ms = []
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step(True)
        ms = (C.c_float * 9)()
        info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        stage_ms.append(list(ms))
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = b200zk.kernel_launches() - launches0
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    b200zk.check(lib.b200zk_msm_profile(0))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- the commitment the timed steps produced, against the known discrete logs of the synthetic SRS:
    # every rank checks its own partial point, rank 0 the folded point (sum of the ranks' dot products)
    msm_check = None
    try:
        torch.cuda.synchronize()
        my_pt = d_pt.cpu().numpy().view(np.uint64).copy()
        s_host = d_scal.cpu().numpy().view(np.uint64).reshape(n, 4)
        my_dot = dot_mod_r(s_host, point_discrete_logs(SEED_P + k, rank * n, n))
        del s_host
        ok_mine = commitment_matches_discrete_logs(my_pt, my_dot)
        if world > 1:
            flags = torch.tensor([1 if ok_mine else 0], dtype=torch.int64, device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            dots = torch.zeros(world * 4, dtype=torch.int64, device=dev)
            mine = torch.from_numpy(np.frombuffer(my_dot.to_bytes(32, "little"), dtype=np.int64).copy()).to(dev)
            dist.all_gather_into_tensor(dots, mine)
            if rank == 0:
                dl = dots.cpu().numpy().view(np.uint64).reshape(world, 4)
                total = sum(int.from_bytes(dl[r].tobytes(), "little") for r in range(world)) % _FR
                folded_ok = commitment_matches_discrete_logs(np.asarray(result_holder["point"], dtype=np.uint64).reshape(12), total)
                msm_check = {"every_rank_partial_point": bool(flags.item()), "folded_point": bool(folded_ok)}
        else:
            msm_check = {"every_rank_partial_point": bool(ok_mine), "folded_point": bool(ok_mine)}
        if msm_check is not None:
            msm_check["how"] = ("bases are [t_i]G with the documented seeded t_i: the point must equal [sum s_i t_i mod r]G, "
                                "computed with exact Python / numpy integers (bench.py, no library or oracle code)")
    except Exception as ex:   # auxiliary: never take the headline line down
        msm_check = {"error": repr(ex)[:200]}

    # ---- roofline of the dominant kernel (bucket accumulation) + the NTT
    info = list(info)
    npairs = int(info[3])
    names = ["hist", "scan", "scatter", "sync", "accumulate", "combine", "reduce", "reduce_combine", "fold"]
    mean_stage = {nm: statistics.mean(s[i] for s in stage_ms) for i, nm in enumerate(names)}
    acc_ms = mean_stage["accumulate"]
    achieved_modmul = msm_modmul_model(npairs) / (acc_ms * 1e-3)
    ntt_ms = statistics.mean(a.elapsed_time(b) for a, b in ntt_ev)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    traffic = {}
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    except Exception:
        pass
    ntt_gbs = ntt_alg_bytes(k) / (ntt_ms * 1e-3) / 1e9
    ntt_modmul = (1 << k) * ntt_modmul_per_element(k) / (ntt_ms * 1e-3)
    roofline = {
        "kernel": "msm_accum_kernel", "bound": "int",
        "achieved": achieved_modmul / 1e9, "peak": peak_modmul / 1e9, "unit": "Gmodmul/s",
        "frac": achieved_modmul / peak_modmul,
        "peak_source": "b200zk_modmul_peak: register-resident Fq Montgomery chains on all SMs, measured in this run",
        "algorithmic_work": "10 modmul per (point, window) pair x pairs per launch",
        "pairs_per_launch": npairs, "kernel_ms": acc_ms,
        "traffic": traffic.get(f"msm_accum_kernel@k{k}"),
        "hbm": {"algorithmic_bytes": 68.0 * npairs, "achieved": 68.0 * npairs / (acc_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": 68.0 * npairs / (acc_ms * 1e-3) / 1e9 / hbm_peak,
                "note": "the same kernel against the HBM roofline (64 B point + 4 B index per pair): not the bound"},
        "note": "modular integer arithmetic on the IMAD pipe (BASELINE.json north_star): bound = int means the "
                "32x32->64 multiplier pipe; HBM traffic is 68 B/pair, far from its roofline",
    }
    roofline_ntt = {
        "kernel": "ntt_pass_kernel (all passes of one transform)", "bound": "hbm",
        "achieved": ntt_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ntt_gbs / hbm_peak,
        "peak_source": hbm_src, "algorithmic_bytes": ntt_alg_bytes(k), "kernel_ms": ntt_ms,
        "int_frac": ntt_modmul / peak_modmul,
        "traffic": traffic.get(f"ntt@k{k}"),
        "note": "254-bit butterflies make the transform integer-bound on B200: int_frac is the fraction of the "
                "measured modmul peak, counting only the multiplications the kernel executes (ntt_modmul_per_element: 12.0 per element at k = 24)",
    }

    # ---- the same MSM through plain best_multiexp (bases per call, no window table)
    with torch.cuda.stream(stream):
        b200zk.check(lib.b200zk_msm_g1_dev_async(vp(d_scal), vp(d_base), n, vp(d_pt), st))
        torch.cuda.synchronize()
        u0 = torch.cuda.Event(enable_timing=True)
        u1 = torch.cuda.Event(enable_timing=True)
        u0.record(stream)
        for _ in range(3):
            b200zk.check(lib.b200zk_msm_g1_dev_async(vp(d_scal), vp(d_base), n, vp(d_pt), st))
        u1.record(stream)
        torch.cuda.synchronize()
    unreg_ms = u0.elapsed_time(u1) / 3

    # ---- e2e: the host-buffer C-ABI calls, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, b200zk, lib, dev, world, rank, k, n, handle, d_scal, d_pt, gathered, omega, barrier)

    # ---- restated CPU baseline on this box's host cores (rank 0, N = 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import c_oracle as co
        co.build()
        threads = co.host_threads()
        ck = args.cpu_k or min(k, 24)
        if ck == k:     # the step's own inputs (the seeded generators agree bit for bit, tests/test_oracle.py)
            cpu_scalars = d_scal.cpu().numpy().view(np.uint64).reshape(n, 4).copy()
            secs, msm_s, fft_s = cpu_step(co, ck, threads, cpu_scalars, h_bases_np)
            del cpu_scalars
            what = f"one whole step at 2^{ck} on the GPU arm's own inputs"
            # the CPU result doubles as a full-size parity check of the commit the GPU arm timed
            from oracle import bn254 as bn
            gpu_pt = d_pt.cpu().numpy().view(np.uint64)
            msm_matches = bn.g1_jacobian_limbs_to_affine(gpu_pt) == bn.g1_jacobian_limbs_to_affine(cpu_step.last_point)
        else:
            secs, msm_s, fft_s = cpu_step(co, ck, threads)
            what = f"one step at 2^{ck}"
        cpu_baseline = {"value": (1 << ck) / secs / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{what} (best_multiexp {msm_s:.2f} s + best_fft {fft_s:.2f} s); "
                                  f"C restatement of the rayon CPU path, {threads} threads",
                        "msm_mpts_per_s": (1 << ck) / msm_s / 1e6, "ntt_melem_per_s": (1 << ck) / fft_s / 1e6}
        if ck == k:
            cpu_baseline["same_commitment_as_gpu"] = bool(msm_matches)
            # ... and of the transform: best_fft of the same 2^k scalars, element for element
            d_chk = d_scal.clone()
            torch.cuda.synchronize()          # the clone runs on torch's stream, the transform on the library's
            b200zk.check(lib.b200zk_ntt_dev(vp(d_chk), n, 1, k, _ptr(omega), None, None))
            torch.cuda.synchronize()
            gpu_fft = d_chk.cpu().numpy().view(np.uint64).reshape(n, 4)
            cpu_baseline["same_transform_as_gpu"] = bool(np.array_equal(gpu_fft, cpu_step.last_fft))
            del d_chk, gpu_fft
            cpu_step.last_fft = None

    # ---- N > 1: ONE best_fft of 2^k sharded over all ranks (four-step, one exchange over NVLink)
    # (auxiliary legs never take the headline line down with them)
    ntt_sharded = None
    if world > 1:
        try:
            ntt_sharded = run_sharded_ntt(torch, dist, lib, b200zk, world, rank, dev, k, ntt_ms)
        except Exception as e:   # noqa: BLE001
            ntt_sharded = {"error": repr(e)}

    # ---- configs[2] stand-in: one create_proof hot path at the RSA-SHA256 circuit shape
    proof_shape = None
    if not args.no_proof_shape:
        try:
            proof_shape = run_proof_shape(torch, dist, world, rank, dev, cpu=(rank == 0 and world == 1 and not args.no_cpu))
        except Exception as e:   # noqa: BLE001
            proof_shape = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (254-bit Montgomery Fr/Fq)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_ntt": roofline_ntt, "cpu_baseline": cpu_baseline,
            "msm_point_verified": msm_check,
            "msm": {"ms": sum(mean_stage.values()), "mpts_per_s": n / (sum(mean_stage.values()) * 1e-3) / 1e6,
                    "window_bits": int(info[1]), "windows": int(info[2]), "stages_ms": mean_stage},
            "ntt": {"ms": ntt_ms, "alg_GBps": ntt_gbs, "melem_per_s": n / (ntt_ms * 1e-3) / 1e6},
            "msm_unregistered": {"ms": unreg_ms, "mpts_per_s": n / (unreg_ms * 1e-3) / 1e6,
                                 "api": "b200zk_msm_g1_dev_async (best_multiexp, bases passed per call)"},
            "srs_registration_s": t_reg, "proof_shape": proof_shape, "ntt_sharded": ntt_sharded,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, b200zk, lib, dev, world, rank, k, n, handle, d_scal, d_pt, gathered, omega, barrier):
    """The step through the host-pointer C ABI, page-locked host polynomials, every copy inside the timed region.

    Headline (`value`): the prover's pattern — `commit_lagrange(p)` then a transform of the *same* polynomial p, as
    create_proof does with every advice / lookup / permutation column (commit_lagrange, then lagrange_to_coeff).  With
    the library's device mirrors on (b200zk_mirror_enable) the commit's upload is the only upload: the transform finds p
    in HBM, runs there and writes the result back.  Every step starts from host data the library has not seen
    (b200zk_mirror_invalidate: new witness values), so a step moves 2^k * 32 B up and 2^k * 32 B + 96 B down.
    `separate_buffers`: round 1's definition (two unrelated host buffers, no mirrors: 2 * 2^k * 32 B up), kept for
    comparison.  With N > 1 the ranks run free (staggered starts, the one-point gather + fold one step behind: see
    below); the in-step schedule of the first half of round 2 is timed beside it (`lockstep`)."""
    from b200zk.api import _ptr
    vp = lambda t: C.c_void_p(t.data_ptr())
    out = np.zeros(12, dtype=np.uint64)

    def gather_fold():
        if world > 1:
            d_pt.copy_(torch.from_numpy(out.view(np.int64)))
            dist.all_gather_into_tensor(gathered, d_pt)
            if rank == 0:
                pts = gathered.cpu().numpy().view(np.uint64).reshape(world, 12)
                b200zk.g1_sum(np.ascontiguousarray(pts))

    def timed(step_fn, steps, warm):
        for _ in range(warm):
            step_fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- headline: one polynomial committed, then transformed (two buffers so that the odd ranks can lag by one)
    polys = [torch.empty(n * 4, dtype=torch.int64, pin_memory=True) for _ in range(2)]
    for p in polys:
        p.copy_(d_scal)
    torch.cuda.synchronize()
    b200zk.check(lib.b200zk_mirror_enable(4 * n * 32))
    state = {"i": 0, "commit_s": 0.0, "transform_s": 0.0, "calls": 0}
    lag = world > 1 and rank % 2 == 1

    def commit(buf):
        t0 = time.perf_counter()
        b200zk.check(lib.b200zk_msm_g1_registered(handle.value, vp(buf), n, _ptr(out)))
        state["commit_s"] += time.perf_counter() - t0

    def transform(buf):
        t0 = time.perf_counter()
        b200zk.check(lib.b200zk_ntt(vp(buf), k, _ptr(omega)))
        state["transform_s"] += time.perf_counter() - t0
        state["calls"] += 1

    def shared_step():
        cur, prev = polys[state["i"] & 1], polys[(state["i"] + 1) & 1]
        state["i"] += 1
        b200zk.check(lib.b200zk_mirror_invalidate(vp(cur), 0))       # new host data for this step's polynomial
        if lag:
            transform(prev)                                           # the polynomial committed one step ago
            commit(cur)
        else:
            commit(cur)
            transform(cur)
        gather_fold()

    warm = max(2, min(args.warmup, 3))
    if world == 1:
        for _ in range(warm):
            shared_step()
        state.update(commit_s=0.0, transform_s=0.0, calls=0)
        dt = timed(shared_step, args.steps, 0)
        order = "commit, then transform"
        lockstep = None
    else:
        # (a) round 2's first schedule, kept for comparison: every rank in step (the gather + fold closes every step),
        # odd ranks one polynomial behind so that half of them upload while the other half downloads
        for _ in range(warm):
            shared_step()
        steps_a = max(2, min(args.steps, 5))
        dt_a = timed(shared_step, steps_a, 0)
        lockstep = {"value": world * n * steps_a / dt_a / 1e6, "ms_per_step": 1e3 * dt_a / steps_a, "steps": steps_a,
                    "what": "all ranks in step: gather + fold of the commitments at the end of every step, odd ranks one "
                            "polynomial behind"}
        # (b) headline: the ranks run free.  The commitments of a step are gathered asynchronously and folded one step
        # later (nothing in a step waits for another rank), and the ranks' starts are staggered (below), so that the eight
        # 0.5 GiB downloads — what the host link cannot take all at once (`pcie_concurrent`) — follow one another
        # instead of coinciding.  Every copy, every gather and every fold is still inside the timed region.
        lag = False
        gbuf = [torch.zeros(world * 12, dtype=torch.int64, device=dev) for _ in range(2)]
        dbuf = [torch.zeros(12, dtype=torch.int64, device=dev) for _ in range(2)]
        pending = []

        def fold_all_but(keep):
            while len(pending) > keep:
                w, i = pending.pop(0)
                w.wait()
                if rank == 0:
                    pts = gbuf[i].cpu().numpy().view(np.uint64).reshape(world, 12)
                    b200zk.g1_sum(np.ascontiguousarray(pts))

        behind = [False]       # odd ranks: transform the polynomial committed one step ago, then commit the next

        def free_step():
            cur, prev = polys[state["i"] & 1], polys[(state["i"] + 1) & 1]
            state["i"] += 1
            b200zk.check(lib.b200zk_mirror_invalidate(vp(cur), 0))
            if behind[0]:
                transform(prev)
                commit(cur)
            else:
                commit(cur)
                transform(cur)
            i = state["i"] & 1
            dbuf[i].copy_(torch.from_numpy(out.view(np.int64)))
            pending.append((dist.all_gather_into_tensor(gbuf[i], dbuf[i], async_op=True), i))
            fold_all_but(1)

        t0 = time.perf_counter()
        for _ in range(2):
            free_step()
        fold_all_but(0)
        torch.cuda.synchronize()
        est = torch.tensor([(time.perf_counter() - t0) / 2], dtype=torch.float64, device=dev)
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
        # Download windows (the last ~10 ms of a transform call) spread over the step at the least cost in idle starts:
        # the odd ranks run one polynomial behind — their window falls a quarter of a step after their neighbour's at no
        # cost — and every further pair of ranks starts 8 ms (or 1/N of a step, if shorter) after the pair before
        behind[0] = rank % 2 == 1
        offset = min(float(est.item()) / world, 0.008) * (rank // 2)
        state.update(commit_s=0.0, transform_s=0.0, calls=0)
        barrier()
        t0 = time.perf_counter()
        time.sleep(offset)
        for _ in range(args.steps):
            free_step()
        fold_all_but(0)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        order = ("ranks run free: odd ranks one polynomial behind, every further pair of ranks starts 8 ms later, the commitments "
                 "of a step are gathered asynchronously and folded one step later; all of it inside the timed region")
    e2e = {"value": world * n * args.steps / dt / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": world * n * 32, "d2h_bytes_per_step": world * (n * 32 + 96),
           "ms_per_step": 1e3 * dt / args.steps,
           "calls_ms_rank0": {"commit_lagrange": 1e3 * state["commit_s"] / max(1, state["calls"]),
                              "best_fft": 1e3 * state["transform_s"] / max(1, state["calls"])},
           "api": "b200zk_msm_g1_registered (ParamsKZG::commit_lagrange: the polynomial from page-locked host memory, SRS "
                  "resident) + b200zk_ntt (best_fft in place on the same host polynomial, found in its device mirror: "
                  "b200zk_mirror_enable; b200zk_mirror_invalidate before every step: the host data is new)",
           "rank_order": order}
    if lockstep is not None:
        e2e["lockstep"] = lockstep
    if world > 1:      # every rank's own call times (the step is the slowest rank's)
        mine = torch.tensor([1e3 * state["commit_s"] / max(1, state["calls"]), 1e3 * state["transform_s"] / max(1, state["calls"])],
                            dtype=torch.float64, device=dev)
        allr = torch.zeros(2 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        e2e["calls_ms_per_rank"] = {"commit_lagrange": [round(float(x), 2) for x in allr[0::2].tolist()],
                                    "best_fft": [round(float(x), 2) for x in allr[1::2].tolist()]}
    # same results as the device-resident calls on the same inputs (rank 0, single GPU)
    if rank == 0 and world == 1:
        try:
            chk = polys[0]
            chk.copy_(d_scal)
            d_chk = d_scal.clone()
            torch.cuda.synchronize()
            b200zk.check(lib.b200zk_mirror_invalidate(vp(chk), 0))
            b200zk.check(lib.b200zk_msm_g1_registered(handle.value, vp(chk), n, _ptr(out)))
            b200zk.check(lib.b200zk_ntt(vp(chk), k, _ptr(omega)))
            b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, 1, n, vp(d_pt), None))
            b200zk.check(lib.b200zk_ntt_dev(vp(d_chk), n, 1, k, _ptr(omega), None, None))
            torch.cuda.synchronize()
            dev_pt = d_pt.cpu().numpy().view(np.uint64).reshape(1, 12)
            same_pt = bool(np.array_equal(b200zk.g1_to_bytes(np.ascontiguousarray(out.reshape(1, 12))),
                                          b200zk.g1_to_bytes(np.ascontiguousarray(dev_pt))))
            e2e["same_results_as_device_resident_calls"] = same_pt and bool(torch.equal(chk, d_chk.cpu()))
            st = (C.c_uint64 * 4)()
            b200zk.check(lib.b200zk_mirror_stats(st))
            e2e["mirror_stats"] = {"hits": int(st[0]), "misses": int(st[1]), "resident_bytes": int(st[2])}
            del d_chk
        except Exception as ex:   # auxiliary: never take the headline line down
            e2e["verify_error"] = str(ex)[:200]
    b200zk.check(lib.b200zk_mirror_enable(0))

    # ---- round 1's definition: two unrelated host buffers, no mirrors (1 GiB up + 0.5 GiB down at k = 24)
    try:
        h_scal, h_ntt = polys
        h_scal.copy_(d_scal); h_ntt.copy_(d_scal)
        torch.cuda.synchronize()

        def separate_step():
            b200zk.check(lib.b200zk_msm_g1_registered(handle.value, vp(h_scal), n, _ptr(out)))
            gather_fold()
            b200zk.check(lib.b200zk_ntt(vp(h_ntt), k, _ptr(omega)))

        steps2 = max(2, min(args.steps, 5))
        dt2 = timed(separate_step, steps2, 1)
        e2e["separate_buffers"] = {"value": world * n * steps2 / dt2 / 1e6, "ms_per_step": 1e3 * dt2 / steps2, "steps": steps2,
                                   "h2d_bytes_per_step": world * n * 64, "d2h_bytes_per_step": world * (n * 32 + 96),
                                   "note": "round 1's e2e: the commit and the transform on two unrelated host buffers, no mirrors"}
    except Exception as ex:
        e2e["separate_buffers"] = {"error": str(ex)[:200]}

    # ---- the host link: copy rates of all ranks at once (512 MiB each way; at N = 1 simply this box's PCIe rates)
    if True:
        try:
            nb = min(n * 32, 1 << 29)
            hbuf, dbuf = polys[0][: nb // 8], torch.empty(nb // 8, dtype=torch.int64, device=dev)

            def rate(fn):
                fn(); torch.cuda.synchronize(); barrier()
                t0 = time.perf_counter()
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                mine = 3 * nb / (time.perf_counter() - t0) / 1e9
                t = torch.tensor([mine], dtype=torch.float64, device=dev)
                lo, total = t.clone(), t.clone()
                if world > 1:
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                    dist.all_reduce(total, op=dist.ReduceOp.SUM)
                return {"min_rank_GBps": float(lo.item()), "sum_GBps": float(total.item())}

            up = lambda: dbuf.copy_(hbuf, non_blocking=True)
            down = lambda: hbuf.copy_(dbuf, non_blocking=True)
            e2e["pcie_concurrent"] = {"all_ranks_h2d": rate(up), "all_ranks_d2h": rate(down),
                                      "even_h2d_odd_d2h": rate(up if rank % 2 == 0 else down),
                                      "bytes_per_copy": nb}
            del dbuf
        except Exception as ex:
            e2e["pcie_concurrent"] = {"error": str(ex)[:200]}

    # informational: the round-1 step on a pageable numpy buffer (what a Rust Vec<Fr> is), and on that buffer after
    # b200zk_host_register (what the shim does once per long-lived buffer)
    if rank == 0 and world == 1:
        try:
            pg_scal = polys[0].numpy().view(np.uint64).reshape(n, 4).copy()
            pg_ntt = pg_scal.copy()

            def host_step():
                b200zk.check(lib.b200zk_msm_g1_registered(handle.value, _ptr(pg_scal), n, _ptr(out)))
                b200zk.check(lib.b200zk_ntt(_ptr(pg_ntt), k, _ptr(omega)))

            def wall(reps=2):
                host_step()
                t0 = time.perf_counter()
                for _ in range(reps):
                    host_step()
                return 1e3 * (time.perf_counter() - t0) / reps

            e2e["separate_buffers"]["pageable_ms_per_step"] = wall()
            with b200zk.pinned(pg_scal), b200zk.pinned(pg_ntt):
                e2e["separate_buffers"]["registered_ms_per_step"] = wall()
            del pg_scal, pg_ntt
        except Exception as ex:   # auxiliary: never take the headline line down
            e2e["pageable_error"] = str(ex)[:200]
    return e2e


def run_sharded_ntt(torch, dist, lib, b200zk, world: int, rank: int, dev, k: int, single_ms: float):
    """One 2^k best_fft over all ranks (strong scaling of a single transform): the transform's first pass on the
    rank's columns stores straight into the peers' row buffers over NVLink (falls back to a NCCL all-to-all when
    peer mappings are unavailable), then the local row transforms."""
    from b200zk import sharding
    from b200zk.api import _ptr as _ptr_np
    vp = lambda t: C.c_void_p(t.data_ptr())
    n = 1 << k
    log_n1 = sharding.four_step_split(k, world)
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    m = n2 // world
    x0 = torch.empty(m * n1 * 4, dtype=torch.int64, device=dev)
    b200zk.check(lib.b200zk_gen_scalars_dev(vp(x0), m * n1, SEED_S + k, rank * m * n1))
    ops = sharding.DeviceFourStep(k, world, rank, dev, mode="auto")
    omega = omega_for(k)
    reps = 5
    xs = [x0.clone() for _ in range(reps + 1)]
    rows_out = sharding.sharded_best_fft(xs[reps], k, omega, ops, world, rank)
    torch.cuda.synchronize()
    # parity, live: the natural-order vector whose column blocks are the ranks' slabs, transformed on this GPU alone,
    # must hold this rank's rows (row i1, position i2  <->  A[i1 + n1 i2])
    full = torch.empty(n * 4, dtype=torch.int64, device=dev)
    slab = torch.empty(m * n1 * 4, dtype=torch.int64, device=dev)
    for r in range(world):
        b200zk.check(lib.b200zk_gen_scalars_dev(vp(slab), m * n1, SEED_S + k, r * m * n1))
        full.view(n1, n2, 4)[:, r * m:(r + 1) * m].copy_(slab.view(n1, m, 4))
    b200zk.check(lib.b200zk_ntt_dev(vp(full), n, 1, k, _ptr_np(fr_limbs(omega)), None, None))
    torch.cuda.synchronize()
    r0, nrows = rank * (n1 // world), n1 // world
    want = full.view(n2, n1, 4)[:, r0:r0 + nrows].permute(1, 0, 2).contiguous().view(-1)
    same = torch.tensor([1 if torch.equal(rows_out.view(-1), want) else 0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    del full, slab, want
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        sharding.sharded_best_fft(xs[i], k, omega, ops, world, rank)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"k": k, "ms": ms, "melem_per_s": n / ms / 1e3, "exchange": ops.mode,
            "same_transform_as_single_gpu": bool(same.item()), "split": f"2^{log_n1} x 2^{k - log_n1}",
            "speedup_vs_one_gpu": single_ms / ms, "exchange_bytes_per_rank": (world - 1) * n * 32 // (world * world),
            "layout": "rank r holds columns j2 in [r m, (r+1) m) in, rows i1 in [r n1/N, (r+1) n1/N) out"}


def run_proof_shape(torch, dist, world: int, rank: int, dev, cpu: bool):
    import b200zk as b200zk_mod
    """The hot-path call sequence of one create_proof at the RSA-SHA256 circuit shape
    (k = 15, extended k = 17; SURVEY.md section 8 d) on synthetic columns, every rank one
    replica.  Returns per-stage milliseconds (median of 3 after one warm-up)."""
    from b200zk.prover_shape import RSA_SHA256, ProverHotPath
    hp = ProverHotPath(RSA_SHA256, sync=torch.cuda.synchronize)
    hp.run()
    runs = [hp.run() for _ in range(3)]
    med = {k: statistics.median(r[k] for r in runs) for k in runs[0]}
    tt = torch.tensor([med["total"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    overlapped = None
    try:       # the same proof with the transforms issued under the commitments (two streams)
        hi, lo = torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev, priority=0)
        hp.run_overlapped(torch, hi, lo)
        overlapped = statistics.median(hp.run_overlapped(torch, hi, lo)["total"] for _ in range(3))
    except Exception as e:   # noqa: BLE001
        overlapped = {"error": repr(e)[:200]}
    # ---- BASELINE.json configs[4] stand-in, executed: a batch of 64 independent proof-shaped workloads, 64 / 8 = 8 per
    # GPU (replicas only: nothing is shared but the read-only SRS / proving key), each with its commitments and
    # transforms on two streams; proofs/s over all GPUs from the slowest rank's wall time
    batch = None
    try:
        per_gpu = 8
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(per_gpu):
            hp.run_overlapped(torch, hi, lo)
        torch.cuda.synchronize()
        bt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        batch = {"proofs": world * per_gpu, "per_gpu": per_gpu, "wall_ms": 1e3 * float(bt.item()),
                 "proofs_per_s": world * per_gpu / float(bt.item()),
                 "what": f"{per_gpu} proof-shaped hot paths back to back on every GPU (64 in all on 8 GPUs), measured, max over ranks"}
    except Exception as e:   # noqa: BLE001
        batch = {"error": repr(e)[:200]}
    sharded = None
    if world > 1:
        # ONE proof over all ranks: columns dealt for commit / iNTT, coefficient columns all-gathered,
        # every rank extends and evaluates its cosets of the extended domain, rank 0 finishes h(X)
        want = hp.h_coeff.to_host().copy()
        want_pts = b200zk_mod.g1_to_bytes(np.ascontiguousarray(hp.points.to_host().reshape(-1, 12)))
        hp.prepare_sharded(world, rank)
        dealt = [list(hp.commit_ranges)]
        for _ in range(3):        # measure, re-deal the commitments by the measured load, repeat
            hp.rebalance(torch, dist, hp.run_sharded(torch, dist))
            dealt.append(list(hp.commit_ranges))
        hp.run_sharded(torch, dist)
        same = True
        if rank == 0:
            got_pts = b200zk_mod.g1_to_bytes(np.ascontiguousarray(hp.points.to_host().reshape(-1, 12)))
            keep = list(range(hp.n_lag)) + [hp.n_lag + 1 + j for j in range(RSA_SHA256.degree - 1)]
            same = bool(np.array_equal(hp.h_coeff.to_host(), want)) and bool(np.array_equal(got_pts[keep], want_pts[keep]))
        sruns = [hp.run_sharded(torch, dist) for _ in range(3)]
        smed = {k: statistics.median(r[k] for r in sruns) for k in sruns[0]}
        st = torch.tensor([smed["total"]], dtype=torch.float64, device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        # the stage every rank spends longest in, and the slowest rank's share of it
        names = [k for k in smed if k != "total"]
        per_rank = torch.tensor([smed[k] for k in names], dtype=torch.float64, device=dev)
        allr = torch.zeros(world * len(names), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, per_rank)
        table = allr.view(world, len(names)).cpu().tolist()
        worst = max(range(len(names)), key=lambda i: max(row[i] for row in table))
        sharded = {"one_proof_over_all_gpus_ms": float(st.item()), "speedup_vs_one_gpu": med["total"] / float(st.item()),
                   "stages_ms_rank0": smed, "stages_ms_max_over_ranks": {nm: max(row[i] for row in table) for i, nm in enumerate(names)},
                   "limiting_stage": names[worst],
                   "commit_columns_per_rank": [e - b for b, e in hp.commit_ranges],
                   "h_coefficients_and_commitments_equal_single_gpu": same,
                   "schedule": "own share to coefficient form + all-gather; coset owners extend and evaluate their cosets; "
                               "commitments dealt by measured load (coset owners and rank 0 commit fewer columns); the evaluated "
                               "cosets travel to rank 0 asynchronously; rank 0 finishes h(X)",
                   "exchange": "all-gather of coefficient columns (NCCL) + asynchronous all-gather of evaluated h cosets + all-gather of commitments"}
    # ---- the same proof through the ABI the Rust patch binds under an *untouched* create_proof: one synchronous
    # host-pointer call per polynomial (page-locked host polynomials), without and with device mirrors
    percall = None
    try:
        hp.run()
        want_h = hp.h_coeff.to_host().copy()
        pts = hp.points.to_host().reshape(-1, 12)
        want_pts = np.concatenate([pts[: hp.n_lag], pts[hp.n_lag + 1: hp.n_lag + 1 + RSA_SHA256.degree - 1]])
        hp.prepare_percall(pinned=True)

        def leg(mirror):
            hp.run_percall(mirror=mirror)
            runs = [hp.run_percall(mirror=mirror) for _ in range(3)]
            same = bool(np.array_equal(hp.h_hcoeff, want_h)) and bool(
                np.array_equal(b200zk_mod.g1_to_bytes(np.ascontiguousarray(hp.h_points)), b200zk_mod.g1_to_bytes(np.ascontiguousarray(want_pts))))
            return {k: statistics.median(r[k] for r in runs) for k in runs[0]}, same

        plain, same_plain = leg(False)
        mirrored, same_mirrored = leg(True)
        percall = {"what": "one synchronous host-pointer C-ABI call per polynomial, as rust/halo2_proofs_patch binds them under an "
                           "unmodified create_proof (host polynomials page-locked); evaluate_h is its replaced body: uploads of "
                           "the coefficient columns / product cosets it is handed, device transforms + kernels, one download",
                   "calls": hp.percall_counts(),
                   "stages_ms": plain, "total_ms": plain["total"], "same_outputs_as_device_resident": same_plain,
                   "with_mirrors": {"stages_ms": mirrored, "total_ms": mirrored["total"], "same_outputs_as_device_resident": same_mirrored,
                                    "mirror_stats": hp.mirror_stats},
                   "vs_device_resident": plain["total"] / med["total"], "with_mirrors_vs_device_resident": mirrored["total"] / med["total"]}
    except Exception as e:   # noqa: BLE001
        percall = {"error": repr(e)[:300]}
    out = {"workload": "synthetic stand-in for BASELINE.json configs[2] (RSA-SHA256 sub-circuit proof): the commit / "
                       "iNTT / coset-NTT / evaluate_h / extended_to_coeff calls of one create_proof on seeded random "
                       "columns of the circuit's shape; the advice columns are uploaded from page-locked host memory and every commitment is downloaded inside the timed region; witness synthesis, transcript and SHPLONK opening excluded",
           "calls": hp.counts(), "stages_ms": med, "hot_path_ms": float(tt.item()),
           "hot_path_overlapped_ms": overlapped, "per_call_host_pointer_abi": percall if (world == 1) else None,
           "proofs_per_s_all_gpus": world / (float(tt.item()) * 1e-3), "parallelism": "replicas only",
           "batch_of_independent_proofs": batch,
           "sharded": sharded}
    hp.close()
    if cpu:
        from oracle import c_oracle as co
        co.build()
        threads = co.host_threads()
        s = RSA_SHA256
        n = 1 << s.k
        c = out["calls"]
        # the real call counts, one call at a time as create_proof issues them, each on its own column
        n_msm = c["commit_lagrange"] + c["commit"]
        cols = co.gen_scalars(SEED_S + s.k, 8 * n).reshape(8, n, 4)
        pts = co.gen_points(SEED_P + s.k, n, threads=threads)
        ext_cols = co.gen_scalars(SEED_S + s.k + 2, 2 * 4 * n).reshape(2, 4 * n, 4)
        w_k, w_ext = fr_limbs(omega_for(s.k)), fr_limbs(omega_for(s.k + 2))
        t0 = time.perf_counter()
        for i in range(n_msm):
            co.best_multiexp(cols[i % 8], pts, threads)
        t1 = time.perf_counter()
        for i in range(c["lagrange_to_coeff"]):
            co.best_fft(cols[i % 8], w_k, s.k, threads)
        t2 = time.perf_counter()
        for i in range(c["coeff_to_extended"] + c["extended_to_coeff"]):
            co.best_fft(ext_cols[i % 2], w_ext, s.k + 2, threads)
        t3 = time.perf_counter()
        out["cpu_baseline"] = {"value": 1e3 * (t3 - t0), "unit": "ms", "cores": threads, "kind": "port",
                               "stages_ms": {"best_multiexp": 1e3 * (t1 - t0), "best_fft_2^k": 1e3 * (t2 - t1),
                                             "best_fft_2^ext_k": 1e3 * (t3 - t2)},
                               "sample": f"every call of the proof timed one by one on the C restatement: {n_msm} best_multiexp at 2^{s.k}, "
                                         f"{c['lagrange_to_coeff']} best_fft at 2^{s.k}, {c['coeff_to_extended'] + c['extended_to_coeff']} at "
                                         f"2^{s.k + 2}; evaluate_h's per-point loop has no C restatement and is not included (lower bound of the CPU hot path)"}
    return out


def sweep(args, torch, b200zk, lib, dev) -> None:
    """k = 16..26 table: MSM Mpts/s and NTT GB/s with inputs resident in HBM (1 GPU)."""
    from b200zk.api import _ptr
    stream = torch.cuda.Stream(device=dev)
    st = C.c_void_p(stream.cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    peak = b200zk.modmul_peak(4096)
    rows = []
    b200zk.check(lib.b200zk_msm_profile(1))
    for k in range(16, 27):
        n = 1 << k
        d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
        d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
        d_pt = torch.zeros(12, dtype=torch.int64, device=dev)
        b200zk.check(lib.b200zk_gen_scalars_dev(vp(d_scal), n, SEED_S + k, 0))
        b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), n, SEED_P + k, 0))
        d_ntt = d_scal.clone()
        omega = fr_limbs(omega_for(k))
        reps = 5 if k <= 22 else 3
        with torch.cuda.stream(stream):
            for _ in range(2):
                b200zk.check(lib.b200zk_msm_g1_dev_async(vp(d_scal), vp(d_base), n, vp(d_pt), st))
                b200zk.check(lib.b200zk_ntt_dev(vp(d_ntt), n, 1, k, _ptr(omega), None, st))
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            for _ in range(reps):
                b200zk.check(lib.b200zk_msm_g1_dev_async(vp(d_scal), vp(d_base), n, vp(d_pt), st))
            e[1].record(stream)
            for _ in range(reps):
                b200zk.check(lib.b200zk_ntt_dev(vp(d_ntt), n, 1, k, _ptr(omega), None, st))
            e[2].record(stream)
            torch.cuda.synchronize()
        msm_ms = e[0].elapsed_time(e[1]) / reps
        ntt_ms = e[1].elapsed_time(e[2]) / reps
        ms = (C.c_float * 9)()
        info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        acc = 10.0 * info[3] / (ms[4] * 1e-3)
        ntt_mm = n * ntt_modmul_per_element(k) / (ntt_ms * 1e-3)
        rows.append({"k": k, "msm_ms": msm_ms, "msm_mpts_per_s": n / msm_ms / 1e3, "c": int(info[1]),
                     "accum_ms": ms[4], "accum_frac_of_modmul_peak": acc / peak,
                     "ntt_ms": ntt_ms, "ntt_alg_GBps": ntt_alg_bytes(k) / ntt_ms / 1e6,
                     "ntt_frac_of_modmul_peak": ntt_mm / peak})
        del d_scal, d_base, d_ntt
    # ---- adversarial scalar distributions (SURVEY.md section 8 d) at k = 22
    k = 22
    n = 1 << k
    d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
    d_pt = torch.zeros(12, dtype=torch.int64, device=dev)
    b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), n, SEED_P + k, 0))
    uniform = torch.empty(n * 4, dtype=torch.int64, device=dev)
    b200zk.check(lib.b200zk_gen_scalars_dev(vp(uniform), n, SEED_S + k, 0))
    minus_one = torch.from_numpy(fr_limbs(-1).view(np.int64)).to(dev).repeat(n)
    gen = torch.Generator(device=dev); gen.manual_seed(7)
    keep = (torch.rand(n, device=dev, generator=gen) < 0.1).to(torch.int64).unsqueeze(1)
    sparse = (uniform.view(n, 4) * keep).reshape(-1).contiguous()
    zero_one = torch.zeros(n, 4, dtype=torch.int64, device=dev)
    zero_one[:, :] = torch.from_numpy(fr_limbs(1).view(np.int64)).to(dev)
    zero_one = (zero_one * (torch.rand(n, device=dev, generator=gen) < 0.5).to(torch.int64).unsqueeze(1)).reshape(-1).contiguous()
    adversarial = {}
    for name, sc in (("uniform", uniform), ("all_r_minus_1", minus_one), ("90pct_zero", sparse), ("bits_0_1", zero_one)):
        with torch.cuda.stream(stream):
            b200zk.check(lib.b200zk_msm_g1_dev_async(vp(sc), vp(d_base), n, vp(d_pt), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                b200zk.check(lib.b200zk_msm_g1_dev_async(vp(sc), vp(d_base), n, vp(d_pt), st))
            e1.record(stream)
            torch.cuda.synchronize()
        adversarial[name] = {"msm_ms": e0.elapsed_time(e1) / 3}
    print(json.dumps({"sweep": rows, "modmul_peak_G_per_s": peak / 1e9,
                      "adversarial_scalars_k22": adversarial}), flush=True)


if __name__ == "__main__":
    main()
