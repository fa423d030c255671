#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few numbers DESIGN.md / bench.py quote (per launch):
    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep > profiles/rN_ncu_full_summary.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "tmem_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_lsu_wavefronts",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum": "smem_tensor_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_ctas",
    "launch__block_size": "block_threads",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem",
    "launch__cluster_size": "cluster_size",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "us": 1.0, "ms": 1e3, "msecond": 1e3,
              "ns": 1e-3, "nsecond": 1e-3}

out = {}
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")], "grid": r[head.index("Grid Size")], "block": r[head.index("Block Size")]}
        for col, name in WANT.items():
            if col in head:
                i = head.index(col)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if name in ("dram_read", "dram_write", "l2_bytes", "dynamic_smem"):
                    d[name + "_bytes"] = v * UNIT_SCALE.get(u, 1.0)
                elif name == "duration":
                    d["duration_us"] = v * UNIT_SCALE.get(u, 1.0)
                else:
                    d[name] = v
        launches.append(d)
    out[path.split("/")[-1]] = launches
json.dump(out, sys.stdout, indent=1)
