"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into the per-launch JSON kept under profiles/.

  python tools/ncu_summarize.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_config3_metrics.json "<source note>"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "ms",
    "dram__bytes_read.sum": "dram_read_GB",
    "dram__bytes_write.sum": "dram_write_GB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_cycles_pct_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_pct_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    if rep.endswith(".csv"):          # already exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X.csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(header)}
    launches = []
    for r in rows[2:]:
        if len(r) < len(header):
            continue
        d = {"id": int(r[col["ID"]]), "kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for m, name in KEEP.items():
            if m not in col:
                continue
            v = r[col[m]].replace(",", "")
            try:
                x = float(v)
            except ValueError:
                continue
            u = units[col[m]]
            if name in ("ms", "dram_read_GB", "dram_write_GB") and u in UNIT_SCALE:
                x *= UNIT_SCALE[u]
            d[name] = round(x, 4)
        launches.append(d)
    json.dump({"source": note, "launches": launches}, open(out, "w"), indent=1)
    print(f"{len(launches)} launches -> {out}")


if __name__ == "__main__":
    main()
