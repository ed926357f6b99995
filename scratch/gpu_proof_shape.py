import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import torch, b200zk
from b200zk.prover_shape import RSA_SHA256, SMALL, ProverHotPath
b200zk.init(0)
for shape in (SMALL, RSA_SHA256):
    hp = ProverHotPath(shape, sync=torch.cuda.synchronize)
    hp.run()
    for _ in range(2):
        print(json.dumps({k: round(v, 3) for k, v in hp.run().items()}), flush=True)
    print(hp.counts(), flush=True)
    hp.close()
