"""Writes tests/golden/bench_tree_H{36,37}.npz: the tables (Mecano depth-first order) of the humanoid trees that bench.py
evaluates, as produced by the host model (MultiBodySystemRandomTools.nextHumanoid, seed bench.HUMANOID_SEED).

The reference arm of the benchmark (`bench.py --impl reference`) and its `cpu_baseline` leg build the CPU oracle from this
file, so that neither imports the product package; the GPU arm checks at run time that the tree it generates equals the
file, i.e. that both arms evaluate the same multi-body system.

    python tests/golden/make_bench_tree.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import bench  # noqa: E402
import mecano_b200 as mb  # noqa: E402

for neck in (1, 2):
    elevator = mb.RigidBody("elevator")
    mb.MultiBodySystemRandomTools.nextHumanoid(bench.HUMANOID_SEED, elevator, neck)
    system = mb.MultiBodySystem.toMultiBodySystemBasics(elevator)
    d = system.describe()
    path = os.path.join(HERE, "bench_tree_H%d.npz" % d["nv"])
    np.savez(path, seed=bench.HUMANOID_SEED, **{k: np.asarray(v) for k, v in d.items()})
    print(path, {k: np.asarray(v).shape for k, v in d.items()})
