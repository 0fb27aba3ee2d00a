"""Shared-memory wavefronts (ideal vs actual) and stall reasons per source line, from an ncu report.
usage: ncu_conflicts.py report.ncu-rep kernel_substring"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; kern = sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
src = list(csv.reader(io.StringIO(out)))
sections = []; cur = None
for r in src:
    if r and r[0] == 'File Path': cur = {'file': r[1], 'rows': []}; sections.append(cur)
    elif r and r[0] == 'Function Name': cur['func'] = r[1]
    elif r and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and r: cur['rows'].append(r)
wf = collections.Counter(); ideal = collections.Counter(); txt = {}
stall = collections.Counter(); stall_line = collections.defaultdict(collections.Counter)
for s in sections:
    if kern not in s.get('func', ''): continue
    h = s['hdr']; iW = h.index('L1 Wavefronts Shared'); iI = h.index('L1 Wavefronts Shared Ideal')
    st = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    f = s['file'].split('/')[-1]
    for r in s['rows']:
        if not r[0]: continue
        try: w = int(r[iW] or 0); i = int(r[iI] or 0)
        except ValueError: continue
        key = (f, int(r[0])); wf[key] += w; ideal[key] += i; txt[key] = r[1].strip()[:90]
        for ci, n in st:
            try: v = int(r[ci] or 0)
            except ValueError: v = 0
            stall[n] += v; stall_line[n][key] += v
tw = sum(wf.values()); ti = sum(ideal.values())
print("shared wavefronts %d, ideal %d (%.1f%% excess)" % (tw, ti, 100.0 * (tw - ti) / max(tw, 1)))
for k, v in wf.most_common(25):
    print("%6.2f%% wf  x%.2f  %s:%d  %s" % (100.0 * v / tw, v / max(ideal[k], 1), k[0], k[1], txt[k]))
ts = sum(stall.values())
print("stall samples:")
for n, v in stall.most_common(8):
    top = ", ".join("%s:%d %.1f%%" % (k[0][:12], k[1], 100.0 * c / ts) for k, c in stall_line[n].most_common(4))
    print("%6.2f%%  %-22s %s" % (100.0 * v / ts, n, top))
