"""Fuzz of the host-side welding (FixedJoints / ignored joints lumped at flatten time, mecano_b200/multibody.py) through the
kernels' per-state code compiled for the host: random trees, random weld sets, both modes, checked against the oracle on the full
tree with the welded joints held still (tests/test_welds.py: check_pair).  CPU only.

    python scripts/fuzz_welds.py [seconds] [first_seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_welds as tw  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    t0 = time.time()
    cases = skipped = 0
    fails = []
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        case = dict(seed=seed, n_joints=int(rng.integers(2, 40)), weld_fraction=float(rng.uniform(0.05, 0.7)), floating=bool(rng.integers(0, 2)),
                    mode=str(rng.choice(["fixed", "ignore"])))
        seed += 1
        try:
            welded, full, held = tw.build_pair(**case)
            if case["mode"] == "ignore":
                held = tw.held_closure(full, held)
            if not held:
                skipped += 1
                continue
            tw.check_pair(welded, full, held, tw.run_emu, seed=seed)
            cases += 1
        except AssertionError as ex:
            fails.append((case, str(ex)[:200]))
    print("cases %d, nothing welded in %d, failures %d, seeds up to %d" % (cases, skipped, len(fails), seed - 1))
    for f in fails[:10]:
        print("FAIL", f)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
