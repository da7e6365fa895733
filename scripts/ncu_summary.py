"""Print the key metrics of an .ncu-rep (first profiled launch) as text; used to write profiles/*.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__block_size', 'launch__grid_size',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum']
want += [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w:88s} {units[i]:16s} {data[0][i]}")
