"""Copy the evidence of one gpu_round.sh visit (gpurun_out/<tag>) into profiles/ and print the headline numbers."""
import csv, json, shutil, sys
tag = sys.argv[1]
src = "gpurun_out/%s/" % tag
d = json.loads([x for x in open(src + "bench.log") if x.startswith("{")][-1])
r = json.loads([x for x in open(src + "bench_ref.log") if x.startswith("{")][-1])
print("value %.0f cells/s  %.1f ms/step  e2e %.0f  cpu_baseline %.0f (%d cores)  reference arm %.0f (%d cores)" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"],
    r["value"], r["cpu_baseline"]["cores"]))
print("roofline fp64 frac %.4f  hbm frac %.3e  kernel_ms %.1f  launches %d" % (
    d["roofline"]["frac"], d["roofline"]["hbm"]["frac"], d["roofline"]["kernel_ms"], d["gpu_launches"]))
rows = list(csv.reader(open(src + "ros_full_raw.csv")))
dd = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
keys += [h for h in rows[0] if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = {k: {"value": dd[k][0], "unit": dd[k][1]} for k in keys if k in dd}
out["_command"] = ("ncu --set full --clock-control none --import-source on -k regex:ros_ -c 1 python bench.py --steps 1 "
                   "--warmup 0 --cells 47360 --no-cpu-baseline")
out["_cells_in_launch"] = 47360
json.dump(out, open("profiles/%s_ros_smem_ncu_full.json" % tag, "w"), indent=1)
f = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd = float(dd["dram__bytes_read.sum"][0]) * f[dd["dram__bytes_read.sum"][1]]
wr = float(dd["dram__bytes_write.sum"][0]) * f[dd["dram__bytes_write.sum"][1]]
json.dump({"dram_bytes_per_cell": (rd + wr) / 47360, "cells_in_capture": 47360, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": "profiles/%s_ros_smem_ncu_full.json (ncu --set full, ros_smem_kernel<fullchem_dims>, 47360 cells)" % tag},
          open("profiles/traffic.json", "w"), indent=1)
for k in ("gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"):
    print(k, dd[k])
print("dram bytes per cell %.0f" % ((rd + wr) / 47360))
for fn in ("launches.csv", "bench.log", "bench_ref.log", "gpu_check.log", "pytest_gpu.log"):
    shutil.copy(src + fn, "profiles/%s_%s" % (tag, fn))
print(open(src + "pytest_gpu.log").read().strip().splitlines()[-5:])
