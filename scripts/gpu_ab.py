"""A/B of builds / run-time switches of the library on the H37 humanoid (2^20 states).  Each variant is `tag[:lib.so][:ENV=V,ENV=V]`
(default: every .so under mecano_b200/variants/); it is timed in its own process for RNEA and ABA (CUDA events, median of 20
launches), the whole series twice; the outputs are hashed so that a change of results is visible.  One JSON line per
(variant, algorithm, round).

    python scripts/gpu_ab.py [variant ...]
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    specs = sys.argv[1:] or [os.path.basename(p)[:-3] + ":" + p for p in sorted(glob.glob(os.path.join(ROOT, "mecano_b200", "variants", "*.so")))]
    for rnd in range(int(os.environ.get("AB_ROUNDS", "2"))):
        for spec in specs:
            parts = spec.split(":")
            env = dict(os.environ, AB_TAG=parts[0])
            for p in parts[1:]:
                if p.endswith(".so"):
                    env["MECANO_B200_LIB"] = os.path.abspath(p)
                elif p:
                    env.update(kv.split("=", 1) for kv in p.split(","))
            for algo in os.environ.get("AB_ALGOS", "rnea,aba").split(","):
                env["AB_ALGO"] = algo
                subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_aba_ab.py"), "child"], env=env, check=False)


if __name__ == "__main__":
    main()
