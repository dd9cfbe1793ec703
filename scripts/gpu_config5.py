"""BASELINE.json config 5: batch-size sweep 1K-16M states (H37) and tree-size sweep 7-101 bodies (thread-per-state against the
body-parallel variant: a warp per state up to 32 bodies, a team of warps beyond), RNEA + ABA + CRBA, on N GPUs of one box.

    python scripts/gpu_config5.py out.jsonl                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/gpu_config5.py out.jsonl                                      # 8 GPUs, one rank each

The batch of B states is split into disjoint slices of B / N states (no collective on the data path; NCCL carries only the
max-over-ranks of the event timings).  Every kernel is timed with CUDA events on its stream (median of `reps` launches after
warm-up); `states_per_s` is B over the slowest rank's time.  The dense mass matrix of more than 2^21 states per GPU
(1,369 x 8 B x 2^21 = 23 GB) is produced in chunks of 2^21 states into one reused buffer (`crba_chunks`), and once whole in the
packed non-zero layout (362 rows)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mecano_b200 as mb  # noqa: E402

CHUNK = 1 << 21


def system(kind, nb):
    e = mb.RigidBody("elevator")
    if kind == "h37":
        mb.MultiBodySystemRandomTools.nextHumanoid(20251017, e, 2)
    else:  # floating base + random one-DoF tree: nb bodies in all
        base = mb.MultiBodySystemRandomTools.nextFloatingBase(100 + nb, e).getSuccessor()
        if nb > 1:
            mb.MultiBodySystemRandomTools.nextOneDoFJointTree(200 + nb, base, nb - 1, 0.0)
    return mb.MultiBodySystem.toMultiBodySystemBasics(e)


def states(s, n, dev, seed):
    nv, nq = s.getNumberOfDoFs(), s.getConfigurationMatrixSize()
    gen = torch.Generator(device=dev).manual_seed(seed)
    q = (torch.rand((nq, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi
    quat = torch.randn((4, n), dtype=torch.float64, device=dev, generator=gen)
    q[0:4] = quat / quat.norm(dim=0, keepdim=True)  # the floating base comes first in every system here
    qd = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    x = torch.rand((nv, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    return q, qd, x


def timed(fn, reps, world):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms[len(ms) // 2]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "-"
    what = (sys.argv[2] if len(sys.argv) > 2 else "batch,tree").split(",")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = open(out_path, "a") if rank == 0 and out_path != "-" else None

    def emit(rec):
        if rank == 0:
            line = json.dumps(rec)
            print(line, flush=True)
            if out:
                out.write(line + "\n")
                out.flush()

    def run_case(sweep, name, s, B, variants):
        n = B // world
        if n < 1:
            return
        nv, nb = s.getNumberOfDoFs(), s.getNumberOfJoints()
        q, qd, x = states(s, n, dev, 1000 + rank)
        res = torch.empty_like(qd)
        for variant in variants:
            calcs = {"rnea": mb.InverseDynamicsCalculator(s, device=local), "aba": mb.ForwardDynamicsCalculator(s, device=local),
                     "crba": mb.CompositeRigidBodyMassMatrixCalculator(s, device=local)}
            try:
                for c in calcs.values():
                    c.setKernelVariant(variant)
            except mb.MecanoB200Error:
                continue
            calcs["rnea"].setGravitationalAcceleration(-9.81)
            calcs["aba"].setGravitationalAcceleration(-9.81)
            reps = 10 if n >= (1 << 16) else 30
            rec = {"sweep": sweep, "tree": name, "bodies": nb, "dofs": nv, "gpus": world, "states": B, "states_per_gpu": n, "variant": variant}
            ms = timed(lambda: calcs["rnea"].compute(q, qd, x, res), reps, world)
            rec["rnea_ms"], rec["rnea_states_per_s"] = ms, B / ms * 1e3
            rec["rnea_kernel"] = calcs["rnea"].kernelInfo(n)["variant"]
            ms = timed(lambda: calcs["aba"].compute(q, qd, x, res), reps, world)
            rec["aba_ms"], rec["aba_states_per_s"] = ms, B / ms * 1e3
            # dense mass matrix, entry-major; beyond CHUNK states per GPU in chunks into one reused buffer
            nc = min(n, CHUNK)
            M = torch.empty((nv * nv, nc), dtype=torch.float64, device=dev)
            chunks = [(o, min(nc, n - o)) for o in range(0, n, nc)]
            # (all matrices of one call share their leading dimension: a chunk of q gets its own contiguous copy, outside the timing)
            qs = [q] if len(chunks) == 1 else [q[:, o:o + m].contiguous() for o, m in chunks]

            def crba_all():
                for (o, m), qc in zip(chunks, qs):
                    calcs["crba"].getMassMatrix(qc, M[:, :m] if m == nc else M[:, :m].contiguous())

            ms = timed(crba_all, max(3, reps // len(chunks)), world)
            rec["crba_ms"], rec["crba_states_per_s"], rec["crba_chunks"] = ms, B / ms * 1e3, len(chunks)
            rec["step_states_per_s"] = B / (rec["rnea_ms"] + rec["aba_ms"] + rec["crba_ms"]) * 1e3
            del M, qs
            if variant != "warp" and sweep == "batch" and n > CHUNK:
                # the packed non-zero layout holds the whole batch
                row, _ = calcs["crba"].getMassMatrixPackedIndex()
                Pk = torch.empty((len(row), n), dtype=torch.float64, device=dev)
                ms = timed(lambda: calcs["crba"].getMassMatrix(q, Pk, packed=True), 5, world)
                rec["crba_packed_ms"], rec["crba_packed_states_per_s"], rec["crba_packed_rows"] = ms, B / ms * 1e3, len(row)
                del Pk
            emit(rec)
        del q, qd, x, res
        torch.cuda.empty_cache()

    if "batch" in what:
        h37 = system("h37", 32)
        for lg in range(10, 25, 2):
            run_case("batch", "h37", h37, 1 << lg, ["auto"])
        # the crossover region, both variants pinned
        for lg in (10, 12, 14, 16):
            run_case("batch_pinned", "h37", h37, (1 << lg) * world, ["thread", "warp"])
    if "tree" in what:
        for nb in (7, 15, 25, 32, 51, 75, 101):
            s = system("tree", nb)
            run_case("tree", "tree%d" % nb, s, (1 << 20) if nb <= 51 else (1 << 19), ["thread"])
            for per_gpu in (256, 1024, 4096, 16384):
                run_case("tree_small", "tree%d" % nb, s, per_gpu * world, ["thread", "warp"])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
