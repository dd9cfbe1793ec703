"""First-contact GPU run: parity of every kernel against the oracle on a few trees, then raw timings on
H37 at 1M states, kernel resource info and the measured roofline denominators.  Writes gpurun_out/first.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_lib as el  # noqa: E402
import oracle_lib as ol  # noqa: E402
import treedesc as td  # noqa: E402

import mecano_b200  # noqa: E402
from mecano_b200 import _capi  # noqa: E402


def engine_for(tree, g):
    d, keep, order = el.tree_desc_c(tree)
    e = mecano_b200.Engine(_capi.TreeDesc.from_buffer_copy(bytes(d)), 0, keepalive=keep)
    e.set_gravity(*g)
    return e, order


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


def main():
    out = {"device": torch.cuda.get_device_name(0)}
    rng = np.random.default_rng(7)
    g = (0.3, -0.2, -9.81)
    dev = torch.device("cuda:0")
    cases = [("chain7", td.chain(rng, 7)), ("tree20p", td.random_tree(rng, 20, prismatic_fraction=0.3)),
             ("float+tree30", td.random_tree(rng, 30, floating=True, com_rotation=True, prismatic_fraction=0.2)),
             ("H37", td.humanoid(rng)), ("tree100", td.random_tree(rng, 100, floating=True))]
    out["parity"] = {}
    for name, t in cases:
        e, order = engine_for(t, g)
        o = ol.Oracle(t, gravity=g)
        n = 1000
        q, qd, qdd, tau = td.random_states(rng, t, n)
        fext = rng.uniform(-1, 1, size=(t.nb, 6, n))
        fo = np.ascontiguousarray(fext.reshape(6 * t.nb, n))
        fe = np.ascontiguousarray(fext[order].reshape(6 * t.nb, n))
        tq, tqd, tqdd, ttau, tf = (torch.from_numpy(x).to(dev) for x in (q, qd, qdd, tau, fe))
        res = {}
        r = torch.empty_like(tqd)
        res["rnea"] = relerr(e.rnea(tq, tqd, tqdd, r).cpu().numpy(), o.rnea_batch(q, qd, qdd))
        res["rnea_fext"] = relerr(e.rnea(tq, tqd, tqdd, r, fext=tf).cpu().numpy(), o.rnea_batch(q, qd, qdd, fo))
        res["aba"] = relerr(e.aba(tq, tqd, ttau, r).cpu().numpy(), o.aba_batch(q, qd, tau))
        res["aba_fext"] = relerr(e.aba(tq, tqd, ttau, r, fext=tf).cpu().numpy(), o.aba_batch(q, qd, tau, fo))
        M = torch.empty((t.nv * t.nv, n), dtype=torch.float64, device=dev)
        res["crba"] = relerr(e.crba(tq, M).cpu().numpy().reshape(t.nv, t.nv, n), o.crba_batch(q))
        Ms = torch.empty((n, t.nv * t.nv), dtype=torch.float64, device=dev)
        res["crba_state_major"] = relerr(e.crba(tq, Ms, layout=1).cpu().numpy().reshape(n, t.nv, t.nv).transpose(1, 2, 0), o.crba_batch(q))
        # host entry point
        th = np.empty_like(qd)
        res["rnea_host"] = relerr(e.rnea_host(q, qd, qdd, th), o.rnea_batch(q, qd, qdd))
        res["info"] = {k: e.kernel_info(a) for k, a in (("rnea", 0), ("aba", 1), ("crba", 2))}
        out["parity"][name] = res
        print(name, {k: v for k, v in res.items() if k != "info"}, flush=True)
        e.close()

    out["fp64_peak_tflops"] = mecano_b200.measure_fp64_peak(0)
    out["hbm_peak_gbs"] = mecano_b200.measure_hbm_peak(0)
    print("peaks", out["fp64_peak_tflops"], out["hbm_peak_gbs"], flush=True)

    # timing on H37, 1M states
    t = td.humanoid(np.random.default_rng(1))
    e, order = engine_for(t, (0, 0, -9.81))
    n = 1 << 20
    gen = torch.Generator(device=dev).manual_seed(0)
    tq = (torch.rand((t.nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    tqd = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    tx = torch.rand((t.nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    r = torch.empty_like(tqd)
    M = torch.empty((t.nv * t.nv, n), dtype=torch.float64, device=dev)
    timing = {}
    for name, fn in (("rnea", lambda: e.rnea(tq, tqd, tx, r)), ("aba", lambda: e.aba(tq, tqd, tx, r)), ("crba", lambda: e.crba(tq, M))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for i in range(10):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
        info = e.kernel_info({"rnea": 0, "aba": 1, "crba": 2}[name])
        timing[name] = {"ms_median": float(np.median(ms)), "ms_min": float(min(ms)), "states_per_s": n / (np.median(ms) * 1e-3),
                        "gbs": info["bytes_per_state"] * n / (np.median(ms) * 1e-3) / 1e9, "info": info}
        print(name, timing[name], flush=True)
    out["timing_h37_1m"] = timing
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "first.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
