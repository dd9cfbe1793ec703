"""Kernel timing sweep over trees and batch sizes (BASELINE.json config 5): prints one JSON line per (tree, algorithm, n).
Usage: python scripts/gpu_sweep.py [trees=h37,chain31f,...] [algos=rnea,aba,crba] [n=1048576,...] [out.jsonl] [variants=thread,warp]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_lib as el  # noqa: E402
import treedesc as td  # noqa: E402

import mecano_b200  # noqa: E402
from mecano_b200 import _capi  # noqa: E402


def make(name, rng):
    if name == "h37":
        return td.humanoid(rng, 2)
    if name == "h36":
        return td.humanoid(rng, 1)
    if name.startswith("chain"):
        fl = name.endswith("f")
        return td.chain(rng, int(name[5:].rstrip("f")), floating=fl)
    if name.startswith("tree"):
        fl = name.endswith("f")
        return td.random_tree(rng, int(name[4:].rstrip("f")), floating=fl)
    raise ValueError(name)


def main():
    trees = (sys.argv[1] if len(sys.argv) > 1 else "h37").split(",")
    algos = (sys.argv[2] if len(sys.argv) > 2 else "rnea,aba,crba").split(",")
    ns = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1048576").split(",")]
    out = open(sys.argv[4], "a") if len(sys.argv) > 4 and sys.argv[4] != "-" else None
    variants = (sys.argv[5] if len(sys.argv) > 5 else "thread").split(",")
    dev = torch.device("cuda:0")
    for name in trees:
        t = make(name, np.random.default_rng(1))
        d, keep, order = el.tree_desc_c(t)
        e = mecano_b200.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
        e.set_gravity(0, 0, -9.81)
        for n in ns:
            if "crba" in algos and t.nv * t.nv * n * 8 > 60e9:
                continue
            gen = torch.Generator(device=dev).manual_seed(0)
            tq = (torch.rand((t.nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
            tqd = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            tx = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            r = torch.empty_like(tqd)
            M = torch.empty((t.nv * t.nv, n), dtype=torch.float64, device=dev) if "crba" in algos else None
            fns = {"rnea": lambda: e.rnea(tq, tqd, tx, r), "aba": lambda: e.aba(tq, tqd, tx, r), "crba": lambda: e.crba(tq, M)}
            for a, variant in [(a, v) for a in algos for v in variants]:
                if variant == "warp" and t.nb > 32:
                    continue
                e.set_variant({"auto": 0, "thread": 1, "warp": 2}[variant])
                if variant != "warp" and os.environ.get("MECANO_B200_SPECIALIZE"):
                    e.specialize([a])
                fn = fns[a]
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                reps = 10
                if os.environ.get("MECANO_B200_SWEEP_GRAPH"):
                    # launch-latency regime: time the kernels alone by replaying a CUDA graph of `per` back-to-back launches
                    per = 20
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        fn()
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr, stream=side):
                            for _ in range(per):
                                fn()
                    torch.cuda.current_stream().wait_stream(side)
                    gr.replay()
                    torch.cuda.synchronize()
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                    ev[0].record()
                    for i in range(reps):
                        gr.replay()
                        ev[i + 1].record()
                    torch.cuda.synchronize()
                    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])) / per
                else:
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                    ev[0].record()
                    for i in range(reps):
                        fn()
                        ev[i + 1].record()
                    torch.cuda.synchronize()
                    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
                info = e.kernel_info({"rnea": 0, "aba": 1, "crba": 2}[a], n)
                line = {"tree": name, "nb": t.nb, "nv": t.nv, "algo": a, "variant": variant, "timing": "graph" if os.environ.get("MECANO_B200_SWEEP_GRAPH") else "launch", "specialized": info["specialized"], "n": n, "ms": ms, "states_per_s": n / (ms * 1e-3),
                        "ns_per_state_body": ms * 1e6 / n / t.nb, "gbs": info["bytes_per_state"] * n / (ms * 1e-3) / 1e9,
                        "block": info["block_threads"], "regs": info["regs_per_thread"], "smem": info["dynamic_smem_bytes"],
                        "blocks_per_sm": info["blocks_per_sm"], "depth": info["max_depth"]}
                print(json.dumps(line), flush=True)
                if out:
                    out.write(json.dumps(line) + "\n")
            del tq, tqd, tx, r, M
        e.close()


if __name__ == "__main__":
    main()
