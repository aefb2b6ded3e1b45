import csv, subprocess, sys
WANT = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size',
 'launch__block_size','launch__waves_per_multiprocessor','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'lts__t_bytes.sum','l1tex__t_bytes.sum','sm__cycles_elapsed.avg','smsp__inst_executed.sum','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__maximum_warps_per_active_cycle_pct']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units=rows[0],rows[1]
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i=hdr.index(w); print(f'  {w:80s} {r[i]} {units[i]}')
        print()
if __name__=='__main__':
    for p in sys.argv[1:]:
        print('==',p); main(p)
