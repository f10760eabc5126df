"""Developer tool: the metrics DESIGN.md / profiles/README.md quote, as a small CSV (metric, unit, one column per launch), from an
ncu report or from its exported raw page.   usage: ncu_subset.py <file.ncu-rep | raw.csv> out.csv"""
import csv, io, subprocess, sys
src, dst = sys.argv[1], sys.argv[2]
text = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
        "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
keep += [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
with open(dst, "w", newline="") as f:
    wr = csv.writer(f)
    names = [r[hdr.index("Kernel Name")].split("(")[0] for r in data]
    wr.writerow(["metric", "unit"] + names)
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            wr.writerow([k, units[i]] + [r[i] for r in data])
print("wrote", dst, len(data), "launches")
