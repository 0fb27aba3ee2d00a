"""Summarise an .ncu-rep: per-kernel headline metrics, stall mix and hottest source lines (development aid)."""
import collections, csv, subprocess, sys, io

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()

def main():
    raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
            'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
    for r in raw[2:]:
        print('----', r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for k in keys:
            if k in hdr:
                i = hdr.index(k); print('   %-62s %s %s' % (k, r[i], units[i]))
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "sass,cuda"]))))
    sections, cur = [], None
    for r in src:
        if r and r[0] == 'File Path':
            cur = {'file': r[1], 'rows': []}; sections.append(cur)
        elif r and r[0] == 'Function Name': cur['func'] = r[1]
        elif r and r[0] == 'Line No': cur['hdr'] = r
        elif cur is not None and r: cur['rows'].append(r)
    byfunc = collections.defaultdict(list)
    for s in sections: byfunc[s['func']].append(s)
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    for f, secs in byfunc.items():
        print('=======', f[:90])
        allrows = []
        for s in secs:
            h = s['hdr']; iInst = h.index('Instructions Executed'); iSamp = h.index('# Samples')
            cols = [(i, x) for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
            for r in s['rows']:
                if not r[0]: continue
                try: inst = int(r[iInst]); samp = int(r[iSamp])
                except ValueError: continue
                st = {x: int(r[i]) for i, x in cols if r[i] not in ('', '0')}
                allrows.append((s['file'].split('/')[-1], r[0], r[1].strip()[:70], inst, samp, st))
        ti = sum(x[3] for x in allrows) or 1; ts = sum(x[4] for x in allrows) or 1
        agg = collections.Counter()
        for x in allrows:
            for k, v in x[5].items(): agg[k] += v
        print('warp-inst (source-attributed)', ti, 'samples', ts)
        print('stall mix:', ' '.join('%s=%.1f%%' % (k.replace('stall_', ''), 100 * v / ts) for k, v in agg.most_common(9)))
        print('-- top lines by stall samples')
        for x in sorted(allrows, key=lambda x: -x[4])[:top]:
            t2 = sorted(x[5].items(), key=lambda kv: -kv[1])[:2]
            print('%5.1f%% samp %5.1f%% inst %-16s:%-4s %-70s %s' % (100 * x[4] / ts, 100 * x[3] / ti, x[0][:16], x[1], x[2], [(k.replace('stall_',''), v) for k, v in t2]))
        print('-- top lines by instructions')
        for x in sorted(allrows, key=lambda x: -x[3])[:top]:
            print('%5.1f%% inst %5.1f%% samp %-16s:%-4s %s' % (100 * x[3] / ti, 100 * x[4] / ts, x[0][:16], x[1], x[2]))

main()
