"""Host pipeline sweep: slots (streams + staging buffers) x chunk size of mecano_b200_step_host on 2^20 H37 states, dense and packed
mass matrix.  One process per configuration (MECANO_B200_HOST_SLOTS / MECANO_B200_HOST_CHUNK_MB are read once)."""
import json
import os
import subprocess
import sys
import time

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
    rng = np.random.default_rng(0)
    step = mb.MultiBodyDynamicsStep(system)
    step.setGravitationalAcceleration(0.0, 0.0, -9.81)
    pin = lambda r: torch.empty((r, N), dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    q, qd, qdd, tau = pin(nq), pin(nv), pin(nv), pin(nv)
    s = mb.MultiBodySystemRandomTools.nextState(rng, system, 4096)
    for dst, src in zip((q, qd, qdd, tau), s):
        dst[:] = np.tile(src, (1, N // 4096))
    to, qo = pin(nv), pin(nv)
    res = {"slots": os.environ.get("MECANO_B200_HOST_SLOTS", "2"), "chunk_mb": os.environ.get("MECANO_B200_HOST_CHUNK_MB", "64")}
    for name, packed in (("packed", True), ("dense", False)):
        M = pin(step.getMassMatrixRows(packed=packed))
        step.compute(q, qd, qdd=qdd, tau=tau, tauOut=to, qddOut=qo, massMatrix=M, packed=packed)
        t0 = time.perf_counter()
        for _ in range(3):
            step.compute(q, qd, qdd=qdd, tau=tau, tauOut=to, qddOut=qo, massMatrix=M, packed=packed)
        res[name + "_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        del M
    print(json.dumps(res), flush=True)


def main():
    for slots, mb_ in (("2", "64"), ("3", "64"), ("4", "64"), ("2", "16"), ("2", "32"), ("2", "128"), ("3", "32"), ("3", "128"), ("4", "32"), ("4", "16")):
        env = dict(os.environ, MECANO_B200_HOST_SLOTS=slots, MECANO_B200_HOST_CHUNK_MB=mb_)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)


if __name__ == "__main__":
    child() if len(sys.argv) > 1 and sys.argv[1] == "child" else main()
