"""Small driver for compute-sanitizer: every entry point once, ragged batch sizes, a humanoid, a branching one-DoF tree, a tree of 70
bodies (the team-per-state kernels) and a tree with spherical / planar joints.

    compute-sanitizer --tool memcheck  python scripts/gpu_memcheck.py
    compute-sanitizer --tool initcheck python scripts/gpu_memcheck.py
    compute-sanitizer --tool racecheck python scripts/gpu_memcheck.py body    # body-parallel kernels only (shared-memory exchange)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mecano_b200 as mb  # noqa: E402

dev = torch.device("cuda:0")
BODY_ONLY = len(sys.argv) > 1 and sys.argv[1] == "body"
for kind in ("humanoid", "tree", "tree70", "joints"):
    e = mb.RigidBody("elevator")
    if kind == "humanoid":
        mb.MultiBodySystemRandomTools.nextHumanoid(3, e, 2)
    elif kind == "tree":
        mb.MultiBodySystemRandomTools.nextOneDoFJointTree(4, e, 20, 0.4)
    elif kind == "tree70":
        base = mb.MultiBodySystemRandomTools.nextFloatingBase(5, e).getSuccessor()
        mb.MultiBodySystemRandomTools.nextOneDoFJointTree(6, base, 69, 0.3)
    else:
        mb.MultiBodySystemRandomTools.nextJointTree(7, e, 18)
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    nb = s.getNumberOfJoints()
    for n in (1, 33, 777) if not BODY_ONLY else (1, 33, 150):
        for variant in ("thread", "warp") if not BODY_ONLY else ("warp",):
            if variant == "warp" and kind == "joints":
                continue  # spherical / planar joints: thread-per-state only
            q, qd, qdd, tau = (torch.from_numpy(x).to(dev) for x in mb.MultiBodySystemRandomTools.nextState(np.random.default_rng(n), s, n))
            fext = torch.rand((6 * nb, n), dtype=torch.float64, device=dev)
            ident = mb.InverseDynamicsCalculator(s).setKernelVariant(variant)
            ident.setGravitationalAcceleration(-9.81)
            ident.compute(q, qd, qdd)
            ident.setExternalWrenches(fext)
            ident.compute(q, qd, qdd)
            fdyn = mb.ForwardDynamicsCalculator(s).setKernelVariant(variant)
            fdyn.compute(q, qd, tau)
            crba = mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant(variant)
            crba.getMassMatrix(q)
            crba.getMassMatrix(q, stateMajor=True)
        if BODY_ONLY:
            torch.cuda.synchronize()
            continue
        full = mb.InverseDynamicsCalculator(s).setComputeByProducts()
        full.compute(q, qd, qdd)
        joints = s.getAllJoints()
        fdyn = mb.ForwardDynamicsCalculator(s)
        fdyn.setJointSourceModes(lambda j: mb.JointSourceMode.ACCELERATION_SOURCE if joints.index(j) % 2 == 0 else None)
        fdyn.compute(q, qd, tau, jointAccelerationInput=qdd)
        cen = mb.CompositeRigidBodyMassMatrixCalculator(s, "centerOfMassFrame")
        cen.getCentroidalMomentumMatrix(q)
        cen.getCentroidalConvectiveTermMatrix(q, qd)
        cen.setEnableCoriolisMatrixCalculation(True)
        cen.getCoriolisMatrix(q, qd)
        integ = mb.MultiBodySystemStateIntegrator(s, 1e-3)
        integ.doubleIntegrateFromAcceleration(q.clone(), qd.clone(), qdd.clone())
        if kind == "humanoid":  # (the fp32 variant exists for humanoid-sized trees)
            for calc, args in ((mb.InverseDynamicsCalculator(s), (q, qd, qdd)), (mb.ForwardDynamicsCalculator(s), (q, qd, tau))):
                calc.setKernelVariant("thread").setPrecision("fp32").compute(*args)
            mb.CompositeRigidBodyMassMatrixCalculator(s).setKernelVariant("thread").setPrecision("fp32").getMassMatrix(q, torch.empty((s.getNumberOfDoFs() ** 2, n), dtype=torch.float64, device=dev))
        # host entry points
        hq, hqd, hqdd = (x.cpu().numpy() for x in (q, qd, qdd))
        full.compute(hq, hqd, hqdd)
        cen.getCentroidalMomentumMatrix(hq)
        cen.getCoriolisMatrix(hq, hqd)
        torch.cuda.synchronize()
print("memcheck driver done")
