import csv, subprocess, sys, io
rep=sys.argv[1]; kern=sys.argv[2]; nw=float(sys.argv[3]); lo=float(sys.argv[4])
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
rows=list(csv.reader(io.StringIO(out)))
hdr=None; cur=None; i=0
for r in rows:
    if r and r[0]=='Kernel Name': cur=r[1]
    elif r and r[0]=='Address': hdr=r
    elif hdr and cur and kern in cur and r and len(r)>=6:
        try: inst=int(r[hdr.index('Instructions Executed')])
        except: continue
        i+=1
        if inst/nw>=lo: print('%5d %6.2f %s' % (i, inst/nw, r[1].strip()[:90]))
