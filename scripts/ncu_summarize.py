"""Condense ncu outputs into the small, tracked summaries kept under profiles/.

  python scripts/ncu_summarize.py launches gpurun_out/X_launches.csv profiles/rNN_launches.md
  python scripts/ncu_summarize.py full gpurun_out/X_prof.ncu-rep profiles/rNN_ncu_full.json

`launches`: the `--metrics gpu__time_duration.sum` list -> per-kernel count / mean / share of the summed device time.
`full`:     a `--set full` report -> the counters DESIGN.md and bench.py quote (DRAM bytes, FP64 pipe, issue, occupancy,
            stall mix), one object per captured launch.
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    m = re.search(r"((?:thread|warp|dfma_chain|copy)_kernel[a-z_0-9]*(?:<[^>]*>)?)", name)
    if m and "at::" not in name:
        return "mb::" + m.group(1).replace("(int)", "").replace("(bool)", "")
    m = re.match(r"void ([A-Za-z_0-9:]+)", name)
    return (m.group(1) if m else name)[:70]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, {"n": 0, "ns": 0.0, "grid": r[8], "block": r[7]})
        a["n"] += 1
        a["ns"] += float(r[-1])
    total = sum(a["ns"] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source: %s\n\n" % src)
        f.write("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n")
        f.write("| kernel | launches | mean us | share of summed device time | grid | block |\n|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %s | %s |\n" % (k, a["n"], a["ns"] / a["n"] / 1e3, 100 * a["ns"] / total, a["grid"], a["block"]))
    print(open(dst).read())


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("sass__inst_executed_local_loads", "local_loads"),
    ("sass__inst_executed_local_stores", "local_stores"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "shared_loads"),
    ("smsp__sass_inst_executed_op_shared_st.sum", "shared_stores"),
    ("smsp__sass_inst_executed_op_global_ld.sum", "global_loads"),
    ("smsp__sass_inst_executed_op_global_st.sum", "global_stores"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dfma_threads"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dmul_threads"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dadd_threads"),
])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for d in data:
        o = OrderedDict(kernel=short(d[idx["Kernel Name"]]))
        for k, nm in WANT.items():
            if k in idx:
                try:
                    o[nm] = float(d[idx[k]])
                except ValueError:
                    o[nm] = d[idx[k]]
                o[nm + "_unit"] = units[idx[k]]
        stalls = {h.split("stalled_")[1]: float(d[i]) for h, i in idx.items() if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
        tot = sum(stalls.values()) or 1.0
        o["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        res.append(o)
    json.dump(res, open(dst, "w"), indent=1)
    for o in res:
        print(json.dumps(o)[:900])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
