"""Launch-configuration sweep: times every compiled configuration (block size, TMEM stack slots; kernels.cu kCfg) of
each algorithm on one tree, each in its own process (MECANO_B200_CFG pins the configuration), and checks that all
configurations produce bit-identical results (the arithmetic does not depend on where the stack lives).
Usage: python scripts/gpu_cfg_sweep.py [tree=h37] [algos=rnea,aba,crba] [cfgs=0,1,2,...] [n=1048576] [out.jsonl]
Child mode: python scripts/gpu_cfg_sweep.py --child tree algo n"""
import json
import os
import subprocess
import sys

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
    out = torch.zeros((t.nv * t.nv if algo == "crba" else t.nv, n), dtype=torch.float64, device=dev)
    ai = {"rnea": 0, "aba": 1, "crba": 2}[algo]

    def run(m):
        # m states of the batch (leading dimension stays n)
        if algo == "rnea":
            e.rnea(tq[:, :m], tqd[:, :m], tx[:, :m], out[:, :m])
        elif algo == "aba":
            e.aba(tq[:, :m], tqd[:, :m], tx[:, :m], out[:, :m])
        else:
            e.crba(tq[:, :m], out[:, :m])

    # ragged tail first: 1000 + 13 states must leave every other column untouched
    m = min(n, 1013)
    run(m)
    torch.cuda.synchronize()
    tail_ok = bool((out[:, m:] == 0).all().item()) if m < n else True
    h_small = int(out[:, :m].contiguous().view(torch.int64).sum().item())
    for _ in range(3):
        run(n)
    torch.cuda.synchronize()
    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(n)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
    info = e.kernel_info(ai)
    h = int(out.view(torch.int64).sum().item())
    print(json.dumps({"tree": tree, "algo": algo, "n": n, "cfg": os.environ.get("MECANO_B200_CFG", ""), "ms": ms, "states_per_s": n / (ms * 1e-3),
                      "block": info["block_threads"], "regs": info["regs_per_thread"], "smem": info["dynamic_smem_bytes"],
                      "local": info["local_bytes_per_thread"], "blocks_per_sm": info["blocks_per_sm"], "hash": h, "hash_small": h_small,
                      "tail_untouched": tail_ok, "finite": bool(torch.isfinite(out).all().item())}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    tree = sys.argv[1] if len(sys.argv) > 1 else "h37"
    algos = (sys.argv[2] if len(sys.argv) > 2 else "rnea,aba,crba").split(",")
    cfgs = (sys.argv[3] if len(sys.argv) > 3 else "0,1,2,3,4,5,6").split(",")
    n = sys.argv[4] if len(sys.argv) > 4 else "1048576"
    out = open(sys.argv[5], "a") if len(sys.argv) > 5 else None
    for a in algos:
        ref = None
        for c in cfgs:
            env = dict(os.environ)
            if c != "auto":
                env["MECANO_B200_CFG"] = "%s=%s" % (a, c)
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", tree, a, n], env=env, capture_output=True, text=True, timeout=120)
                line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else json.dumps({"algo": a, "cfg": c, "error": p.stderr[-400:]})
            except subprocess.TimeoutExpired:
                line = json.dumps({"algo": a, "cfg": c, "error": "timeout"})
            d = json.loads(line)
            if "hash" in d:
                ref = ref or (d["hash"], d["hash_small"])
                d["same_as_first"] = (d["hash"], d["hash_small"]) == ref
                line = json.dumps(d)
            print(line, flush=True)
            if out:
                out.write(line + "\n")
                out.flush()


if __name__ == "__main__":
    main()
