"""Summarise an .ncu-rep: key throughput metrics and the top stall sites.  usage: python tools/ncu_summary.py file.ncu-rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__cluster_size', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum', 'smsp__inst_executed.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:85s} {rows[1][i]:>10s} {[r[i] for r in rows[2:]]}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be concatenated; take the first block
h = None; data = []
for r in rows:
    if r and r[0] == 'Address': 
        if h is not None: break
        h = r; continue
    if h is not None and len(r) == len(h): data.append(r)
si, so = h.index('# Samples'), h.index('Source')
stalls = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
tot = sum(int(r[si] or 0) for r in data)
agg = {}
for r in data:
    for i in stalls:
        agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i] or 0)
print('total samples', tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:ntop]:
    st = {h[i][6:]: int(r[i] or 0) for i in stalls if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}% {r[so][:95]:95s} {st}")
