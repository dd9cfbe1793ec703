"""One step = RNEA + ABA + CRBA on the same batch: sequential on one stream vs the three calculators on three streams with the
ABA / RNEA persistent grids capped (mecano_b200_set_grid_limit).  Prints one JSON line per setting."""
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
q, qd, qdd, tau = (torch.from_numpy(x).to(dev) for x in mb.MultiBodySystemRandomTools.nextState(np.random.default_rng(0), s, n))
nv = s.getNumberOfDoFs()
ident, fdyn, crba = mb.InverseDynamicsCalculator(s), mb.ForwardDynamicsCalculator(s), mb.CompositeRigidBodyMassMatrixCalculator(s)
for c in (ident, fdyn):
    c.setGravitationalAcceleration(-9.81)
o_tau, o_qdd = torch.empty((nv, n), dtype=torch.float64, device=dev), torch.empty((nv, n), dtype=torch.float64, device=dev)
M = torch.empty((nv * nv, n), dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sequential():
    ident.compute(q, qd, qdd, o_tau)
    fdyn.compute(q, qd, tau, o_qdd)
    crba.getMassMatrix(q, M)


def overlapped(order):
    main = torch.cuda.current_stream()
    fork = torch.cuda.Event()
    fork.record(main)
    streams = {"aba": main, "crba": s1, "rnea": s2}
    for name in order:
        st = streams[name]
        if st is not main:
            st.wait_event(fork)
        with torch.cuda.stream(st):
            if name == "aba":
                fdyn.compute(q, qd, tau, o_qdd)
            elif name == "crba":
                crba.getMassMatrix(q, M)
            else:
                ident.compute(q, qd, qdd, o_tau)
    for st in (s1, s2):
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


sequential()
ref = (o_tau.clone(), o_qdd.clone(), M.clone())
print(json.dumps({"mode": "sequential", "ms": timed(sequential)}), flush=True)
for aba_sms in (148, 132, 120, 110, 100, 90, 80, 64):
    for rnea_sms in (148, 48, 32):
        for order in (("aba", "crba", "rnea"), ("aba", "rnea", "crba")):
            fdyn.setGridLimit(aba_sms if aba_sms < 148 else 0)
            ident.setGridLimit(rnea_sms if rnea_sms < 148 else 0)
            o_tau.zero_(); o_qdd.zero_(); M.zero_()
            ms = timed(lambda: overlapped(order))
            same = bool(torch.equal(o_tau, ref[0]) and torch.equal(o_qdd, ref[1]) and torch.equal(M, ref[2]))
            print(json.dumps({"mode": "overlapped", "aba_sms": aba_sms, "rnea_sms": rnea_sms, "order": "+".join(order), "ms": ms, "same": same}), flush=True)
