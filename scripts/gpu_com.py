"""Timing of the centre-of-mass-only launch against the full by-product launch and the convective term (H37, 2^20 states)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mecano_b200 as mb  # noqa: E402
import numpy as np  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    n = 1 << 20
    system = bench.build_system(2)
    rng = np.random.default_rng(1)
    q, qd, _, _ = mb.MultiBodySystemRandomTools.nextState(rng, system, n)
    q, qd = torch.from_numpy(q).cuda(), torch.from_numpy(qd).cuda()
    cen = mb.CompositeRigidBodyMassMatrixCalculator(system, "centerOfMassFrame")
    out = {"cmm_full_ms": timed(lambda: cen.getCentroidalMomentumMatrix(q))}
    com_full = cen.getCenterOfMass().clone()
    out["com_only_ms"] = timed(lambda: cen.getCenterOfMass(q))
    out["bit_identical"] = bool(torch.equal(cen.getCenterOfMass(), com_full))
    out["convective_ms"] = timed(lambda: cen.getCentroidalConvectiveTermMatrix(q, qd))
    out["convective_reuse_ms"] = timed(lambda: cen.getCentroidalConvectiveTermMatrix(q, qd, reuseCenterOfMass=True))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
