"""Tiny driver for ncu: runs each algorithm a few times on the H37 humanoid (args: n_states, reps, algos)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_lib as el  # noqa: E402
import treedesc as td  # noqa: E402

import mecano_b200  # noqa: E402
from mecano_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
algos = sys.argv[3].split(",") if len(sys.argv) > 3 else ["rnea", "aba", "crba"]
t = td.humanoid(np.random.default_rng(1))
d, keep, order = el.tree_desc_c(t)
e = mecano_b200.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
e.set_gravity(0, 0, -9.81)
if os.environ.get("MECANO_B200_SPECIALIZE"):
    e.specialize([a for a in algos if a != "crba"], force=True)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
tq = (torch.rand((t.nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
tqd = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
tx = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
r = torch.empty_like(tqd)
M = torch.empty((t.nv * t.nv, n), dtype=torch.float64, device=dev) if "crba" in algos else None
for _ in range(reps):
    if "rnea" in algos:
        e.rnea(tq, tqd, tx, r)
    if "aba" in algos:
        e.aba(tq, tqd, tx, r)
    if "crba" in algos:
        e.crba(tq, M)
torch.cuda.synchronize()
print("ok", n, reps, algos)
