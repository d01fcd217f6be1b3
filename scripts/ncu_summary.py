"""Extract the headline metrics of every kernel in an .ncu-rep (ncu --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]
tensor_cols = [h for h in hdr if "tensor" in h and h not in want]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    for w in tensor_cols[:12]:
        i = hdr.index(w)
        if r[i] not in ("", "0", "n/a"):
            print(f"{w} = {r[i]} {units[i]}")
    print("---")
