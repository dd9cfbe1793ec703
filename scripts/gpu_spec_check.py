"""Tree-specialised kernels vs the generic kernels on one tree: agreement (relative to the generic result), kernel times,
JIT time.  Each specialised configuration runs in its own process (MECANO_B200_SPEC_CFG pins block size : TMEM slots).
Usage: python scripts/gpu_spec_check.py [tree=h37] [algo=cfg;cfg,...  e.g. rnea=512:32;256:0,aba=256:0] [n=1048576] [out.jsonl]
Child: python scripts/gpu_spec_check.py --child tree algo n"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(tree, algo, n):
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import emu_lib as el
    import gpu_sweep

    import mecano_b200
    from mecano_b200 import _capi

    t = gpu_sweep.make(tree, np.random.default_rng(1))
    d, keep, order = el.tree_desc_c(t)
    e = mecano_b200.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
    e.set_gravity(0, 0, -9.81)
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    tq = (torch.rand((t.nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    tqd = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    tx = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    rows = t.nv * t.nv if algo == "crba" else t.nv
    ai = {"rnea": 0, "aba": 1, "crba": 2}[algo]

    def run(o, m=n):
        if algo == "rnea":
            e.rnea(tq[:, :m], tqd[:, :m], tx[:, :m], o[:, :m])
        elif algo == "aba":
            e.aba(tq[:, :m], tqd[:, :m], tx[:, :m], o[:, :m])
        else:
            e.crba(tq[:, :m], o[:, :m])

    def timeit(o):
        for _ in range(3):
            run(o)
        torch.cuda.synchronize()
        reps = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            run(o)
            ev[i + 1].record()
        torch.cuda.synchronize()
        return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))

    ref = torch.zeros((rows, n), dtype=torch.float64, device=dev)
    ms0 = timeit(ref)
    i0 = e.kernel_info(ai)
    t0 = time.time()
    e.specialize([algo], force=True)
    jit_s = time.time() - t0
    out = torch.zeros((rows, n), dtype=torch.float64, device=dev)
    m = min(n, 1013)
    run(out, m)  # ragged tail: everything beyond m must stay untouched
    torch.cuda.synchronize()
    tail_ok = bool((out[:, m:] == 0).all().item()) if m < n else True
    ms1 = timeit(out)
    i1 = e.kernel_info(ai)
    scale = max(1.0, float(ref.abs().max().item()))
    err = float((out - ref).abs().max().item()) / scale
    print(json.dumps({"tree": tree, "algo": algo, "n": n, "spec_cfg": os.environ.get("MECANO_B200_SPEC_CFG", "default"), "generic_ms": ms0, "spec_ms": ms1,
                      "speedup": ms0 / ms1, "states_per_s": n / (ms1 * 1e-3), "rel_diff_vs_generic": err, "tail_untouched": tail_ok, "jit_s": jit_s,
                      "specialized": i1["specialized"], "block": i1["block_threads"], "tm": i1["tmem_stack_slots"], "regs": i1["regs_per_thread"],
                      "local": i1["local_bytes_per_thread"], "smem": i1["dynamic_smem_bytes"], "blocks_per_sm": i1["blocks_per_sm"],
                      "generic_block": i0["block_threads"], "generic_regs": i0["regs_per_thread"]}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    tree = sys.argv[1] if len(sys.argv) > 1 else "h37"
    spec = sys.argv[2] if len(sys.argv) > 2 else "rnea=default,aba=default"
    n = sys.argv[3] if len(sys.argv) > 3 else "1048576"
    out = open(sys.argv[4], "a") if len(sys.argv) > 4 else None
    for item in spec.split(","):
        algo, cfgs = item.split("=")
        for c in cfgs.split(";"):
            env = dict(os.environ)
            if c != "default":
                env["MECANO_B200_SPEC_CFG"] = "%s=%s" % (algo, c)
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", tree, algo, n], env=env, capture_output=True, text=True, timeout=300)
                line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else json.dumps({"algo": algo, "cfg": c, "error": p.stderr[-600:]})
            except subprocess.TimeoutExpired:
                line = json.dumps({"algo": algo, "cfg": c, "error": "timeout"})
            print(line, flush=True)
            if out:
                out.write(line + "\n")
                out.flush()


if __name__ == "__main__":
    main()
