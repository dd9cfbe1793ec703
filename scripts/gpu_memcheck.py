"""Small driver for compute-sanitizer: every entry point once, ragged batch sizes, a humanoid and a branching one-DoF tree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mecano_b200 as mb  # noqa: E402

dev = torch.device("cuda:0")
for kind in ("humanoid", "tree"):
    e = mb.RigidBody("elevator")
    if kind == "humanoid":
        mb.MultiBodySystemRandomTools.nextHumanoid(3, e, 2)
    else:
        mb.MultiBodySystemRandomTools.nextOneDoFJointTree(4, e, 20, 0.4)
    s = mb.MultiBodySystem.toMultiBodySystemBasics(e)
    nb = s.getNumberOfJoints()
    for n in (1, 33, 777):
        for variant in ("thread", "warp"):
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
        if kind == "humanoid":
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
