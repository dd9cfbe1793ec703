"""Launch-configuration sweep of the Coriolis-matrix kernel (MB_CORIOLIS): MECANO_B200_CFG="cor=K" pins kCfg[K] (kernels.cu)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mecano_b200 as mb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
e = mb.RigidBody("elevator")
mb.MultiBodySystemRandomTools.nextHumanoid(20251017, e, 2)
s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
dev = torch.device("cuda:0")
q, qd, _, _ = mb.MultiBodySystemRandomTools.nextState(np.random.default_rng(0), s, n)
tq, tqd = torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev)
nv = s.getNumberOfDoFs()
ref = None
for cfg in [-1] + list(range(15)):
    if cfg >= 0:
        os.environ["MECANO_B200_CFG"] = "cor=%d" % cfg
    try:
        c = mb.CompositeRigidBodyMassMatrixCalculator(s)
    except Exception as ex:  # the configuration does not fit this tree
        print(json.dumps({"cfg": cfg, "error": str(ex)[:80]}))
        continue
    c.setEnableCoriolisMatrixCalculation(True)
    info = c._engine.kernel_info(3, n)
    for _ in range(2):
        C = c.getCoriolisMatrix(tq, tqd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        C = c.getCoriolisMatrix(tq, tqd)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = C.clone()
    print(json.dumps({"cfg": cfg, "ms": e0.elapsed_time(e1) / 5, "block": info["block_threads"], "regs": info["regs_per_thread"],
                      "local": info["local_bytes_per_thread"], "bps": info["blocks_per_sm"], "smem": info["dynamic_smem_bytes"],
                      "same": bool(torch.equal(C, ref))}), flush=True)
    del c
