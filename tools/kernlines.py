import collections, csv, subprocess, sys, io
rep=sys.argv[1]; kern=sys.argv[2]; fname=sys.argv[3]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
src=list(csv.reader(io.StringIO(out)))
sections=[];cur=None
for r in src:
    if r and r[0]=='File Path': cur={'file':r[1],'rows':[]}; sections.append(cur)
    elif r and r[0]=='Function Name': cur['func']=r[1]
    elif r and r[0]=='Line No': cur['hdr']=r
    elif cur is not None and r: cur['rows'].append(r)
tot=0
rows=[]
allinst=0
for s in sections:
    if kern not in s['func']: continue
    h=s['hdr']; iI=h.index('Instructions Executed')
    for r in s['rows']:
        if not r[0]: continue
        try: inst=int(r[iI])
        except: continue
        allinst+=inst
        if s['file'].endswith(fname): rows.append((int(r[0]), inst, r[1].strip()[:100]))
nwarps=float(sys.argv[4])
for ln,inst,txt in sorted(rows):
    if inst/nwarps>=1.5: print('%4d %6.1f  %s' % (ln, inst/nwarps, txt))
print('total per warp', allinst/nwarps)
