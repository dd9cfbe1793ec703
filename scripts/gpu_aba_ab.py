"""A/B of the ABA pass-three record handling on the H37 humanoid (2^20 states): order of pass three (reversed = stack
discipline over the workspace / forward) x discard of the L2 lines after the read.  Each configuration runs in its own process
(the switches are read once); the joint accelerations of all configurations must agree bit for bit.  Prints one JSON line
per configuration; `--ncu` makes each child run a single launch for an `ncu --metrics dram__bytes_*` wrapper.

    python scripts/gpu_aba_ab.py            # timing
    python scripts/gpu_aba_ab.py child      # (internal)
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 1 << 20


def child():
    import numpy as np
    import torch

    import bench
    import mecano_b200 as mb

    system = bench.build_system(2)
    nv, nq = system.getNumberOfDoFs(), system.getConfigurationMatrixSize()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1)
    q = (torch.rand((nq, N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    quat = torch.randn((4, N), dtype=torch.float64, device=dev, generator=gen)
    q[0:4] = quat / quat.norm(dim=0, keepdim=True)
    qd = torch.rand((nv, N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    tau = torch.rand((nv, N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    out = torch.empty_like(qd)
    algo = os.environ.get("AB_ALGO", "aba")
    if algo == "aba":
        calc = mb.ForwardDynamicsCalculator(system)
    else:
        calc = mb.InverseDynamicsCalculator(system)
    calc.setGravitationalAcceleration(0.0, 0.0, -9.81)
    reps = int(os.environ.get("AB_REPS", "20"))
    for _ in range(3 if reps > 1 else 0):
        calc.compute(q, qd, tau, out)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        calc.compute(q, qd, tau, out)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"algo": algo, "forward": os.environ.get("MECANO_B200_ABA_P3_FORWARD", ""), "discard": os.environ.get("MECANO_B200_ABA_DISCARD", ""),
                      "cfg": os.environ.get("MECANO_B200_CFG", ""), "tag": os.environ.get("AB_TAG", ""), "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "sha": digest, "info": calc.kernelInfo(N)["regs_per_thread"]}))


def main():
    for fwd in ("1", ""):
        for disc in ("0", "1"):
            env = dict(os.environ)
            env.pop("MECANO_B200_ABA_P3_FORWARD", None)
            if fwd:
                env["MECANO_B200_ABA_P3_FORWARD"] = "1"
            env["MECANO_B200_ABA_DISCARD"] = disc
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)


if __name__ == "__main__":
    child() if len(sys.argv) > 1 and sys.argv[1] == "child" else main()
