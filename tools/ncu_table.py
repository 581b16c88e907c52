"""One block of key metrics per launch of an `ncu --set full` report (any kernel).
usage (no GPU needed): ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv [out.txt] [every]
`every` = keep launch k of each group of `every` consecutive launches of the same kernel (default: the last of a run)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Grid Size', 'Block Size', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__cluster_size', 'sm__cycles_elapsed.avg.per_second']
idx = {h: i for i, h in enumerate(hdr)}
ki = idx['Kernel Name']
out = []
# keep the LAST launch of every run of identical kernel names (warm instruction cache, same arguments)
keep = [k for k in range(len(data)) if k + 1 == len(data) or data[k + 1][ki] != data[k][ki]]
for k in keep:
    row = data[k]
    out.append("-- " + row[ki][:150])
    for w in want:
        if w in idx and row[idx[w]] not in ("", "n/a"):
            out.append("  %-70s %s %s" % (w, row[idx[w]], units[idx[w]]))
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
