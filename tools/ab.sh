# usage: tools/ab.sh "ENV=val ..." ...   one quick_time run per argument (empty string = defaults)
for v in "$@"; do
  echo "== ${v:-defaults}"
  env $v timeout 300 python tools/quick_time.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    if d['generic']==0: print(d['config'], round(d['path_ms']*1000,1), round(d['json_ms']*1000,1), round(d['gbps']))
"
done
