"""Key metrics + top stall reasons of every kernel in an ncu raw CSV.  python tools/ncu_detail.py raw.csv [src.csv]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
for r in rows[2:]:
    print(r[idx['Kernel Name']][:70])
    for k in keys:
        if k in idx:
            print('    %-75s %s %s' % (k, r[idx[k]], rows[1][idx[k]]))
    st = [(float(r[idx[h]].replace(',', '') or 0), h) for h in hdr
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for v, h in sorted(st, reverse=True)[:8]:
        print('      stall %.2f %s' % (v, h.split('stalled_')[1].split('_per')[0]))
if len(sys.argv) > 2:
    src = list(csv.reader(open(sys.argv[2])))
    h = src[0]; ix = {c: i for i, c in enumerate(h)}
    samp = [c for c in h if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)' or 'Sampling' in c]
    col = ix.get('Warp Stall Sampling (All Samples)') or ix.get('# Samples')
    if col is None:
        print('sampling columns:', samp); sys.exit()
    tot = 0; lines = []
    for r in src[1:]:
        try: v = float(r[col].replace(',', ''))
        except Exception: continue
        tot += v; lines.append((v, r))
    print('top source/SASS lines by samples (total %d):' % tot)
    for v, r in sorted(lines, key=lambda x: -x[0])[:25]:
        print('   %5.1f%%  %s' % (100 * v / max(tot, 1), ' | '.join(r[ix[c]] for c in h[:3] if c in ix)[:150]))
