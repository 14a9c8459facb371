"""Summarise an .ncu-rep (captured with `ncu --set full --import-source on`) into a small markdown file under
profiles/.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_step_kernels.md "title" [traffic.json n_points]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_static', 'static smem/block'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads per warp instruction (warp efficiency)'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle / issue'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle / issue'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio_throttle / issue'),
]


def ncu(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = ncu(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ['# %s' % title, '', 'Source: `%s` (ncu --set full --clock-control none --import-source on; per-launch values, '
             'cold-ish caches, serialised replays -- compare shares and ratios, not absolute times).' % rep.split('/')[-1], '']
    name_i = hdr.index('Kernel Name')
    for r in data:
        lines += ['## `%s`' % r[name_i].split('(')[0], '', '| metric | value |', '|---|---|']
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                lines.append('| %s | %s %s |' % (label, r[i], units[i]))
        lines.append('')
    # SASS-level stall samples of the first kernel in the report
    src = ncu(rep, 'source')
    if len(src) > 2:
        kname = src[0][1].split('(')[0] if len(src[0]) > 1 else ''
        h = src[1]
        try:
            ia, ins, smp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
            body = [r for r in src[2:] if len(r) > smp]
            tot = sum(int(r[smp] or 0) for r in body) or 1
            by_op = Counter()
            for r in body:
                toks = r[ia].split()
                op = (toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')).split('.')[0]
                by_op[op] += int(r[smp] or 0)
            lines += ['## Stall samples by SASS opcode (`%s`)' % kname, '',
                      ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in by_op.most_common(10)), '',
                      'Hottest SASS lines (samples, executions, instruction):', '']
            for r in sorted(body, key=lambda r: -int(r[smp] or 0))[:8]:
                lines.append('- %s samples, %s executions: `%s`' % (r[smp], r[ins], r[ia].strip()[:100]))
            lines.append('')
        except ValueError:
            pass
    open(dst, 'w').write('\n'.join(lines))
    print('wrote', dst)
    if len(sys.argv) > 5:
        # per-launch DRAM traffic by C-ABI entry point -> profiles/traffic.json (bench.py fills roofline.traffic from it)
        import json
        entry = {'knn_thread_kernel': 'dc_knn', 'knn_record_kernel': 'dc_knn_recorded', 'step_points_kernel': 'dc_step_points', 'step_forward_kernel': 'dc_step_forward',
                 'step_forward_scatter': 'dc_step_forward_scatter',
                 'step_backward_gather_kernel': 'dc_step_backward', 'step_backward_scatter_kernel': 'dc_step_backward_scatter',
                 'step_chain_kernel': 'dc_step_chain'}
        rd, wr, du = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tscale = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}
        out = {'source': '%s (ncu --set full, timed NVTX region of bench.py)' % rep.split('/')[-1], 'n_points': int(sys.argv[5]), 'kernels': {}}
        for r in data:
            base = r[name_i].split('(')[0].replace('void ', '').split('<')[0].strip()
            targs = r[name_i].split('(')[0].split('<')[1] if '<' in r[name_i].split('(')[0] else ''
            if base == 'step_forward_kernel' and ('true' in targs or targs.replace(' ', '').endswith(',1>')):
                base = 'step_forward_scatter'        # SCAT = true: forward + float32 backward scatter in one kernel
            if base in entry and entry[base] not in out['kernels']:
                out['kernels'][entry[base]] = {'kernel': r[name_i].split('(')[0], 'dram_read_bytes': float(r[rd]) * scale[units[rd]],
                                               'dram_write_bytes': float(r[wr]) * scale[units[wr]],
                                               'duration_s_under_ncu': float(r[du]) * tscale[units[du]]}
        json.dump(out, open(sys.argv[4], 'w'), indent=1)
        print('wrote', sys.argv[4])


if __name__ == '__main__':
    main()
