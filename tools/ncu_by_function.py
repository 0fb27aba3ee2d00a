import collections, csv, subprocess, sys, io, re
rep=sys.argv[1]; kern=sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
src=list(csv.reader(io.StringIO(out)))
sections=[];cur=None
for r in src:
    if r and r[0]=='File Path': cur={'file':r[1],'rows':[]}; sections.append(cur)
    elif r and r[0]=='Function Name': cur['func']=r[1]
    elif r and r[0]=='Line No': cur['hdr']=r
    elif cur is not None and r: cur['rows'].append(r)
# map lines to enclosing function by scanning source files
def func_map(path):
    m={}; name='?'
    try: lines=open(path).read().split('\n')
    except: return m
    for i,l in enumerate(lines,1):
        mm=re.match(r'^(?:template.*\n)?(?:RG_HD|__device__ __forceinline__|__global__|inline|static)\b.*?(\w+)\s*\(',l)
        if mm and not l.startswith(' '): name=mm.group(1)
        mm2=re.match(r'^struct (\w+)',l)
        if mm2: name=mm2.group(1)
        mm3=re.match(r'^\s+RG_HD \w[\w\s\*]* (\w+)\(',l)
        if mm3: name=name.split('::')[0]+'::'+mm3.group(1)
        m[i]=name
    return m
tot=0; agg=collections.Counter(); samp=collections.Counter()
for s in sections:
    if kern not in s['func']: continue
    fm=func_map(s['file'].replace('/root/repo/',''))
    h=s['hdr']; iI=h.index('Instructions Executed'); iS=h.index('# Samples')
    for r in s['rows']:
        if not r[0]: continue
        try: inst=int(r[iI]); sm=int(r[iS])
        except: continue
        f=s['file'].split('/')[-1]+':'+fm.get(int(r[0]),'?')
        agg[f]+=inst; samp[f]+=sm; tot+=inst
ts=sum(samp.values())
for k,v in agg.most_common(40):
    print('%5.1f%% inst %5.1f%% samp  %s' % (100*v/tot, 100*samp[k]/ts, k))
