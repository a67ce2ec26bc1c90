"""Print selected metrics of an ncu report (development aid)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu', 'launch__shared_mem_per_block_dynamic']
for vals in rows[2:]:
    print('==', vals[hdr.index('Kernel Name')], 'grid', vals[hdr.index('Grid Size')])
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) for w in want) and 'Triage' not in h and '(Not Issued)' not in h:
            try:
                if float(v.replace(',', '')) == 0: continue
            except ValueError: pass
            print(f"  {h} [{u}] = {v}")
