"""Exchange files for the JVM cross-check (baseline/java/MecanoHarness.java; format documented there).

  python scripts/java_exchange.py write h37 4096 exchange.bin [seed]     # system: a7 | h37 | h36 | tree<N>
  ... run `java MecanoHarness dump exchange.bin results.bin` on a machine with a JDK + the Mecano jars ...
  python scripts/java_exchange.py compare exchange.bin results.bin [--gpu]

`compare` checks Mecano's own results against the C oracle (and with --gpu against the CUDA kernels through the calculator
API) with the error idiom of the reference's tests (ForwardDynamicsCalculatorTest.java:1099-1107) and prints one JSON line;
the north-star bound is 1e-9 relative.  Nothing here runs in the test suites: there is no JVM in the build image.
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MAGIC = 0x4D423258
GRAVITY = (0.0, 0.0, -9.81)


def build(kind, seed):
    import mecano_b200 as mb

    e = mb.RigidBody("elevator")
    if kind == "a7":
        mb.MultiBodySystemRandomTools.nextRevoluteJointChain(seed, e, 7)
    elif kind in ("h37", "h36"):
        mb.MultiBodySystemRandomTools.nextHumanoid(seed, e, 2 if kind == "h37" else 1)
    elif kind.startswith("tree"):
        base = mb.MultiBodySystemRandomTools.nextFloatingBase(seed + 1000, e).getSuccessor()
        mb.MultiBodySystemRandomTools.nextOneDoFJointTree(seed, base, int(kind[4:]), 0.3)
    else:
        raise SystemExit("unknown system " + kind)
    return mb.MultiBodySystem.toMultiBodySystemBasics(e)


def write(kind, n, path, seed=1):
    import mecano_b200 as mb

    s = build(kind, seed)
    d = s.describe()
    q, qd, qdd, tau = mb.MultiBodySystemRandomTools.nextState(np.random.default_rng(seed), s, n)
    with open(path, "wb") as f:
        f.write(struct.pack("<5i3d", MAGIC, d["nb"], d["nv"], d["nq"], n, *GRAVITY))
        for i in range(d["nb"]):
            f.write(struct.pack("<2i", int(d["jtype"][i]), int(d["parent"][i])))
            for key in ("axis", "off_R", "off_p", "com_R", "com_p", "J"):
                f.write(np.ascontiguousarray(d[key][i], dtype="<f8").tobytes())
            f.write(struct.pack("<d", float(d["mass"][i])))
        for a in (q, qd, qdd, tau):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
    print(json.dumps({"written": path, "system": kind, "n_bodies": int(d["nb"]), "n_dofs": int(d["nv"]), "states": n}))


def read(path):
    import treedesc as td

    raw = open(path, "rb").read()
    magic, nb, nv, nq, n = struct.unpack_from("<5i", raw, 0)
    assert magic == MAGIC, "not an exchange file"
    g = struct.unpack_from("<3d", raw, 20)
    off = 44
    f = {k: [] for k in ("jtype", "parent", "axis", "off_R", "off_p", "com_R", "com_p", "J", "mass")}
    for _ in range(nb):
        jt, par = struct.unpack_from("<2i", raw, off)
        off += 8
        v = np.frombuffer(raw, "<f8", 37, off)
        off += 37 * 8
        f["jtype"].append(jt); f["parent"].append(par); f["axis"].append(v[0:3]); f["off_R"].append(v[3:12].reshape(3, 3))
        f["off_p"].append(v[12:15]); f["com_R"].append(v[15:24].reshape(3, 3)); f["com_p"].append(v[24:27]); f["J"].append(v[27:36].reshape(3, 3))
        f["mass"].append(v[36])
    dof, cfg, a, b = [], [], 0, 0
    for jt in f["jtype"]:
        dof.append(a); cfg.append(b)
        a += 6 if jt == 2 else 1
        b += 7 if jt == 2 else 1
    t = td.TreeDesc(nb=nb, nv=nv, nq=nq, parent=np.array(f["parent"], np.int32), jtype=np.array(f["jtype"], np.int32), axis=np.array(f["axis"]),
                    off_R=np.array(f["off_R"]), off_p=np.array(f["off_p"]), com_R=np.array(f["com_R"]), com_p=np.array(f["com_p"]), J=np.array(f["J"]),
                    mass=np.array(f["mass"]), dof_off=np.array(dof, np.int32), cfg_off=np.array(cfg, np.int32)).contiguous()
    mats = []
    for rows in (nq, nv, nv, nv):
        mats.append(np.frombuffer(raw, "<f8", rows * n, off).reshape(rows, n).copy())
        off += rows * n * 8
    return t, g, mats


def err(actual, expected):
    return float(np.max(np.abs(actual - expected)) / max(1.0, np.max(np.abs(expected))))


def compare(exchange, results, gpu=False):
    import oracle_lib as ol

    t, g, (q, qd, qdd, tau) = read(exchange)
    n, nv = q.shape[1], t.nv
    raw = np.fromfile(results, "<f8")
    assert raw.size == (2 * nv + nv * nv) * n, "results.bin does not match exchange.bin"
    j_tau, j_qdd, j_M = raw[:nv * n].reshape(nv, n), raw[nv * n:2 * nv * n].reshape(nv, n), raw[2 * nv * n:].reshape(nv, nv, n)
    o = ol.Oracle(t, gravity=g)
    out = {"states": n, "n_dofs": nv, "tolerance": 1e-9,
           "oracle_vs_jvm": {"rnea": err(o.rnea_batch(q, qd, qdd), j_tau), "aba": err(o.aba_batch(q, qd, tau), j_qdd), "crba": err(o.crba_batch(q), j_M)}}
    if gpu:
        import torch

        import emu_lib as el
        import mecano_b200
        from mecano_b200 import _capi

        d, keep, _ = el.tree_desc_c(t)  # the tables of the file, level-ordered, straight into the C ABI
        e = mecano_b200.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
        e.set_gravity(*g)
        dev = torch.device("cuda:0")
        tq, tqd, tqdd, ttau = (torch.from_numpy(a).to(dev) for a in (q, qd, qdd, tau))
        r = torch.empty_like(tqd)
        M = torch.empty((nv * nv, n), dtype=torch.float64, device=dev)
        out["gpu_vs_jvm"] = {"rnea": err(e.rnea(tq, tqd, tqdd, r).cpu().numpy(), j_tau), "aba": err(e.aba(tq, tqd, ttau, r).cpu().numpy(), j_qdd),
                             "crba": err(e.crba(tq, M).cpu().numpy().reshape(nv, nv, n), j_M)}
    out["pass"] = all(v <= 1e-9 for k in ("oracle_vs_jvm", "gpu_vs_jvm") for v in out.get(k, {}).values())
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "write":
        write(sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 1)
    elif len(sys.argv) >= 4 and sys.argv[1] == "compare":
        compare(sys.argv[2], sys.argv[3], "--gpu" in sys.argv)
    else:
        print(__doc__)
