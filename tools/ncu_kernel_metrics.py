"""Pick the metrics that DESIGN.md / bench.py quote out of an .ncu-rep (one kernel launch per row of `--page raw`).

    python tools/ncu_kernel_metrics.py REPORT.ncu-rep [--kernel REGEX] [--pairs N] > profiles/<name>.json

Prints one JSON object: per matching launch the kernel name, duration, DRAM bytes read / written, issue-slot, ALU / FMA pipe and
LSU wavefront utilisation, registers, occupancy limits and shared-memory bank conflicts."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=".")
    ap.add_argument("--pairs", type=int, default=0, help="pairs scored by the launch (adds dram_bytes_per_pair)")
    ap.add_argument("--what", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ki = names.index("Kernel Name")
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names) or not re.search(a.kernel, r[ki]):
            continue
        m = {}
        for w in WANT:
            if w in names:
                j = names.index(w)
                v = r[j].replace(",", "")
                try:
                    m[w] = float(v)
                except ValueError:
                    m[w] = v
                if units[j]:
                    m[w + ":unit"] = units[j]
        e = dict(kernel=r[ki].split("(")[0], metrics=m)
        if a.pairs and "dram__bytes_read.sum" in m:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = m["dram__bytes_read.sum"] * scale.get(m.get("dram__bytes_read.sum:unit", "byte"), 1)
            wr = m["dram__bytes_write.sum"] * scale.get(m.get("dram__bytes_write.sum:unit", "byte"), 1)
            e["pairs_in_launch"] = a.pairs
            e["dram_bytes_read"] = rd; e["dram_bytes_write"] = wr
            e["dram_bytes_per_pair"] = (rd + wr) / a.pairs
        out.append(e)
    print(json.dumps(dict(what=a.what, report=a.report, launches=out), indent=1))


if __name__ == "__main__":
    main()
