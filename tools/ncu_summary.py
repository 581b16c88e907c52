"""Key metrics of the launches in an `ncu --set full` report of tools/ncu_gemm.py (second launch of each case).
usage (no GPU needed): ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv out.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__cluster_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'smsp__cycles_active.avg']
idx = {h: i for i, h in enumerate(hdr)}
names = ["fwd 3x3 256>256 full-image (M=6080), BN128, cluster split-K 2",
         "fwd 1x1 256>1024 +res +out_r full-image, BN64",
         "fwd 1x1 1024>256 full-image, BN128, cluster split-K 2",
         "dgrad 3x3 crops (M=4608), BN128, cluster split-K 2",
         "wgrad 3x3 full-image, BN128, split-K 4 (RED.ADD)",
         "fwd 3x3 512>256 decoder M=18432, BN256"]
out = ["== tcgen05 implicit-GEMM kernel (TMA producers, round-1 final build): 6 representative launches (tools/ncu_gemm.py)",
       "   source: gpurun_out/r1_ncu_gemm.ncu-rep  (ncu --set full --clock-control none --import-source on); second launch of each case"]
for k, row in enumerate(data):
    if k % 2 == 0:
        continue
    out.append("-- " + names[k // 2])
    for w in want:
        if w in idx:
            out.append("  %-70s %s %s" % (w, row[idx[w]], units[idx[w]]))
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if "tensor" in l or l.startswith("--") or "time_duration" in l))
