#!/usr/bin/env python
"""profiles/ncu_constants.json from .ncu-rep captures of tools/ncu_targets.py: per kernel (the LAST profiled launch of each
name) DRAM bytes per launch, executed instructions and FP64 instructions / FLOP per work unit, pipe utilisations.
bench.py reads the file for `roofline.traffic` and `roofline.executed` instead of carrying literals.

usage: python tools/ncu_constants.py <label> <rep> <units-json> [<rep> <units-json> ...]
   units-json: what tools/ncu_targets.py printed for that capture, e.g. '{"c2": {"path_steps": 1008000000}}'
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_constants.json")
FAMILY = {"heston_fast2_kernel": ("c2", "c2_64"), "heston_f32_kernel": ("c2_f32",), "lsm_backward_kernel": ("c3",), "lsm_paths_kernel": ("c3",),
          "bk_integral_sorted_kernel": ("c4",), "heston_tangent_kernel": ("c5",)}


def fnum(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def to_bytes(val, unit):
    v = fnum(val)
    if v is None:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    label = sys.argv[1]
    consts = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for rep, units_json in zip(sys.argv[2::2], sys.argv[3::2]):
        units = json.loads(units_json)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, unit_row, data = rows[0], rows[1], rows[2:]
        col = {k: i for i, k in enumerate(hdr)}
        for r in data:
            name = re.sub(r"^void ", "", r[col["Kernel Name"]])
            name = re.sub(r"\(.*$", "", name).replace("hh::", "")
            base = name.split("<")[0]
            fam = FAMILY.get(base)
            if not fam:
                continue
            get = lambda k: fnum(r[col[k]]) if k in col else None
            gb = lambda k: to_bytes(r[col[k]], unit_row[col[k]]) if k in col else None
            # which target of this capture the launch belongs to: R64 instantiations are the last-but-one template argument
            tgt = None
            for t in fam:
                if t in units:
                    if base == "heston_fast2_kernel":
                        r64 = re.search(r",\s*(\d),\s*\d>$", name)
                        is64 = bool(r64 and r64.group(1) == "1")
                        if (t == "c2_64") != is64:
                            continue
                    tgt = t
            if tgt is None:
                continue
            work = list(units[tgt].values())[0]
            work_name = list(units[tgt].keys())[0]
            # --set full carries the FP64 opcode counters as rates (thread instructions per elapsed cycle, summed over the
            # SM sub-partitions): totals = rate x elapsed cycles
            cyc = get("smsp__cycles_elapsed.avg") or get("sm__cycles_elapsed.avg") or 0.0

            def total(op):
                v = get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum")
                if v is not None:
                    return v
                return (get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") or 0.0) * cyc
            dadd, dmul, dfma = total("dadd"), total("dmul"), total("dfma")
            inst = get("smsp__inst_executed.sum") or 0.0
            e = {"source": f"profiles/{label} (ncu --set full, {os.path.basename(rep)})", "work_units": work, "work_unit": work_name,
                 "gpu_time_ms": (get("gpu__time_duration.sum") or 0.0) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(
                     unit_row[col["gpu__time_duration.sum"]], 1.0),
                 "dram_bytes_per_launch": (gb("dram__bytes_read.sum") or 0.0) + (gb("dram__bytes_write.sum") or 0.0),
                 "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
                 "instr_per_unit": inst * 32 / work, "fp64_instr_per_unit": (dadd + dmul + dfma) / work,
                 "flop_per_unit": (dadd + dmul + 2 * dfma) / work,
                 "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "fp64_pipe_pct": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                 "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                 "fma_pipe_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                 "xu_pipe_pct": get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                 "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "registers_per_thread": get("launch__registers_per_thread"),
                 "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "thread_inst_per_warp_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio")}
            if work_name == "path_steps":  # the names bench.py reads
                e["instr_per_path_step"], e["fp64_instr_per_path_step"], e["flop_per_path_step"] = (
                    e["instr_per_unit"], e["fp64_instr_per_unit"], e["flop_per_unit"])
            consts[name] = e
    json.dump(consts, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, sorted(consts))


if __name__ == "__main__":
    main()
