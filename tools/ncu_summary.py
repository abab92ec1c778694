#!/usr/bin/env python
"""Summarise an `ncu --set full` report of tools/profile_pass.py into profiles/:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep c2 r01 [--append]

writes profiles/<round>_ncu_summary.md (one block per kernel: duration, DRAM traffic, registers,
issue / pipe utilisation, top stall reasons) and updates profiles/ncu_traffic.json
({workload: {kernel: dram bytes per launch}}), which bench.py reports as `roofline.traffic`.
Runs here (no GPU): it only reads the report with `ncu -i ... --page raw --csv`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"\(.*", "", name).strip()
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"^apap::", "", name)


def main():
    rep, workload, rnd = sys.argv[1], sys.argv[2], sys.argv[3]
    append = len(sys.argv) > 4 and sys.argv[4] == "--append"     # a second capture (more kernels) of the same round
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, out, traffic = set(), [], {}
    for r in body:
        name = short(r[idx["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        out.append(f"### `{name}`\n")
        out.append("| metric | value |\n|---|---|")
        dram = 0.0
        for key, label in KEEP:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                val, unit = r[idx[key]], units[idx[key]]
                out.append(f"| {label} | {val} {unit} |")
                if key.startswith("dram__bytes"):
                    dram += float(val.replace(",", "")) * UNIT_BYTES.get(unit, 1)
        stalls = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i].replace(",", "")))
                  for h, i in idx.items()
                  if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued") and r[i] not in ("", "n/a")]
        tot = sum(v for _, v in stalls) or 1.0
        top = ", ".join(f"{h} {100 * v / tot:.0f}%" for h, v in sorted(stalls, key=lambda x: -x[1])[:5])
        out.append(f"| top stall reasons (sampled) | {top} |\n")
        traffic[name] = int(dram)
    os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
    md = os.path.join(REPO, "profiles", f"{rnd}_ncu_summary.md")
    with open(md, "a" if append else "w") as f:
        f.write(("\n" if append else "") + f"# ncu --set full --clock-control none, `tools/profile_pass.py {workload}` ({os.path.basename(rep)})\n\n"
                "Per-launch values of the first captured launch of each kernel (cold caches, serialised by the "
                "profiler: compare shares, not absolutes).\n\n" + "\n".join(out) + "\n")
    tj = os.path.join(REPO, "profiles", "ncu_traffic.json")
    allt = json.load(open(tj)) if os.path.exists(tj) else {}
    if append:
        allt.setdefault(workload, {}).update(traffic)
    else:
        allt[workload] = traffic                   # the capture replaces the workload's table: no stale kernels of older rounds
    with open(tj, "w") as f:
        json.dump(allt, f, indent=1, sort_keys=True)
    print("wrote", md, "and", tj, traffic)


if __name__ == "__main__":
    main()
