"""Generates tests/golden/oracle_golden.npz: small seeded input/output vectors produced by the C oracle
(oracle/mecano_oracle.c).  These are REGRESSION fixtures for the oracle and the CUDA path; they are not outputs of
the Java reference (which cannot run here: no JDK).  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import treedesc as td  # noqa: E402


def main():
    rng = np.random.default_rng(424242)
    out = {}
    for name, t in (("A7", td.chain(rng, 7)), ("tree12", td.random_tree(rng, 12, floating=True, prismatic_fraction=0.3, com_rotation=True)),
                    ("H37", td.humanoid(rng, 2))):
        g = (0.0, 0.0, -9.81)
        o = ol.Oracle(t, gravity=g)
        n = 6
        q, qd, qdd, tau = td.random_states(rng, t, n)
        fext = np.ascontiguousarray(rng.uniform(-1, 1, size=(6 * t.nb, n)))
        for f in ("parent", "jtype", "axis", "off_R", "off_p", "com_R", "com_p", "J", "mass", "dof_off", "cfg_off"):
            out["%s/%s" % (name, f)] = getattr(t, f)
        out[name + "/dims"] = np.array([t.nb, t.nv, t.nq])
        out[name + "/gravity"] = np.array(g)
        for k, v in (("q", q), ("qd", qd), ("qdd", qdd), ("tau", tau), ("fext", fext)):
            out["%s/%s" % (name, k)] = v
        out[name + "/rnea"] = o.rnea_batch(q, qd, qdd, fext)
        out[name + "/aba"] = o.aba_batch(q, qd, tau, fext)
        out[name + "/crba"] = o.crba_batch(q)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote oracle_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
