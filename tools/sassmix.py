import collections, csv, subprocess, sys, io, re
rep=sys.argv[1]; kern=sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
rows=list(csv.reader(io.StringIO(out)))
hdr=None; cur=None; ops=collections.Counter(); tot=0; lines=[]; seen=set()
for r in rows:
    if r and r[0]=='Kernel Name': cur=r[1]
    elif r and r[0]=='Address': hdr=r
    elif hdr and cur and kern in cur and r and len(r)>=6:
        try: inst=int(r[hdr.index('Instructions Executed')])
        except: continue
        sass=r[1]
        if (cur, r[0]) in seen: continue        # the page can list a kernel's SASS more than once
        seen.add((cur, r[0]))
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', sass)
        op=m.group(2) if m else sass[:10]
        ops[op]+=inst; tot+=inst; lines.append((inst,sass.strip()))
nw=float(sys.argv[3]) if len(sys.argv)>3 else 1
print('total', tot, 'per warp', tot/nw)
for k,v in ops.most_common(36): print('%6.2f%% %7.1f  %s' % (100*v/tot, v/nw, k))
