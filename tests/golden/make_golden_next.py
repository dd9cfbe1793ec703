"""Generates tests/golden/oracle_golden_next.npz: regression vectors of the oracle for the SURVEY 8f rows (RNEA by-products,
forward dynamics with joint source modes, centroidal momentum matrix / convective term, Coriolis matrix) on the trees and
states of oracle_golden.npz.  Like that file these come from the C oracle (oracle/mecano_oracle.c), NOT from the Java
reference (no JDK here).  Re-run:  python tests/golden/make_golden_next.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import treedesc as td  # noqa: E402

FIELDS = ("parent", "jtype", "axis", "off_R", "off_p", "com_R", "com_p", "J", "mass", "dof_off", "cfg_off")


def load_cases():
    data = np.load(os.path.join(HERE, "oracle_golden.npz"))
    for name in sorted({k.split("/")[0] for k in data.files}):
        dims = data[name + "/dims"]
        t = td.TreeDesc(**{f: data["%s/%s" % (name, f)] for f in FIELDS}, nb=int(dims[0]), nv=int(dims[1]), nq=int(dims[2])).contiguous()
        yield name, t, tuple(data[name + "/gravity"]), {k: data["%s/%s" % (name, k)] for k in ("q", "qd", "qdd", "tau", "fext")}


def evaluate(t, g, s, locked):
    """All next-row quantities of every state of one case, stacked along the last axis."""
    o, o0 = ol.Oracle(t, gravity=g), ol.Oracle(t, gravity=(0.0, 0.0, 0.0))
    n = s["q"].shape[1]
    out = {k: [] for k in ("acc", "wr", "qdd_src", "tau_src", "cmm_world", "cmm_com", "com", "conv_world", "conv_com", "coriolis")}
    for k in range(n):
        q, qd, qdd, tau = (s[x][:, k] for x in ("q", "qd", "qdd", "tau"))
        fext = np.ascontiguousarray(s["fext"][:, k].reshape(t.nb, 6))
        _, acc, wr = o.rnea_full(q, qd, qdd, fext)
        a, tq = o.aba_sources(q, qd, tau, qdd, locked, fext)
        _, A0, com, mass = o0.crba_centroidal(q, 0)
        _, A1, _, _ = o0.crba_centroidal(q, 1)
        _, C = o0.coriolis(q, qd)
        for key, v in (("acc", acc), ("wr", wr), ("qdd_src", a), ("tau_src", tq), ("cmm_world", A0), ("cmm_com", A1), ("com", np.append(com, mass)),
                       ("conv_world", o0.centroidal_convective_term(q, qd, 0)), ("conv_com", o0.centroidal_convective_term(q, qd, 1)), ("coriolis", C)):
            out[key].append(v)
    return {k: np.stack(v, axis=-1) for k, v in out.items()}


def main():
    rng = np.random.default_rng(515151)
    out = {}
    for name, t, g, s in load_cases():
        locked = (rng.uniform(size=t.nb) < 0.4).astype(np.int32)
        locked[rng.integers(t.nb)] = 1
        out[name + "/accel_source"] = locked
        for k, v in evaluate(t, g, s, locked).items():
            out["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "oracle_golden_next.npz"), **out)
    print("wrote oracle_golden_next.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
