"""Key metrics + top stalled SASS lines from an .ncu-rep (uses the local ncu CLI)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_uniform.sum', 'launch__grid_size', 'launch__block_size']
for r in data:
    print('===', r[hdr.index('Kernel Name')][:70])
    for i, h in enumerate(hdr):
        if h in want:
            print(f'   {h} [{units[i]}] {r[i]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(r[isamp].rjust(6), r[iex].rjust(9), r[isrc][:84].ljust(84), [x for x in st if x[0]])
