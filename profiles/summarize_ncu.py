"""Summarise an .ncu-rep (read on the GPU-less build box): python profiles/summarize_ncu.py rep.ncu-rep > out.txt
Keeps the metrics the roofline needs (B200_PROFILING.md): duration, DRAM bytes, tensor-pipe activity, registers."""
import csv
import subprocess
import sys

PAT = ("gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum",
       "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__mem_tensor_cycles_active.avg",
       "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct",
       "sm__cycles_elapsed.avg ", "sm__cycles_active.avg ", "lts__t_bytes.sum ", "smsp__inst_executed.sum ",
       "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum ",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_static",
       "launch__occupancy_limit", "sm__inst_executed_pipe_tensor_op_hmma.sum ")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "| block", d.get("Block Size"))
        for i, k in enumerate(hdr):
            kk = k + " "
            if any(p in kk for p in PAT) and ".peak_sustained" not in k and ".per_second" not in k and \
                    ".max" not in k.replace("sm__cycles_elapsed.max", "") and ".min" not in k and ".sum.pct" not in k:
                print(f"  {k} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
