"""Round-2 probe: the RSA-shaped proof through the device-handle pipeline, the per-call host-pointer ABI and the
same with mirrors.  python scratch/r2_percall.py"""
import json, statistics, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
import numpy as np, torch
import b200zk
from b200zk.prover_shape import RSA_SHA256, ProverHotPath

b200zk.init(0)
hp = ProverHotPath(RSA_SHA256, sync=torch.cuda.synchronize)
hp.run()
runs = [hp.run() for _ in range(3)]
med = {k: round(statistics.median(r[k] for r in runs), 3) for k in runs[0]}
print("device-resident", json.dumps(med), flush=True)
want_h = hp.h_coeff.to_host().copy()
hp.prepare_percall(pinned=True)
for mirror in (False, True):
    hp.run_percall(mirror=mirror)
    rs = [hp.run_percall(mirror=mirror) for _ in range(3)]
    m = {k: round(statistics.median(r[k] for r in rs), 3) for k in rs[0]}
    print("percall mirror=%s" % mirror, json.dumps(m), "same_h", bool(np.array_equal(hp.h_hcoeff, want_h)),
          getattr(hp, "mirror_stats", None) if mirror else "", flush=True)
hp.close()
