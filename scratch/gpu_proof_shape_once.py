import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import torch, b200zk
from b200zk.prover_shape import RSA_SHA256, ProverHotPath
b200zk.init(0)
hp = ProverHotPath(RSA_SHA256, sync=torch.cuda.synchronize)
hp.run()
print(json.dumps({k: round(v, 3) for k, v in hp.run().items()}), flush=True)
