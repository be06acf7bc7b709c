"""Key per-launch metrics of an ncu report: python tools/ncu_key.py file.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum']
for r in rows[2:]:
    print('---', r[h.index('Kernel Name')][:60])
    for k in keys:
        if k in h:
            print('  %-70s %s' % (k, r[h.index(k)]))
    st = []
    for i, k in enumerate(h):
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            st.append((float(r[i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print('  stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:9]))
