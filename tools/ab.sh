# usage: tools/ab.sh LIB ...   one quick_time run per library (path relative to the repo root; "" = the product build)
for v in "$@"; do
  echo "== ${v:-libregk.so}"
  REGK_LIB=${v:+$PWD/$v} timeout 300 python tools/quick_time.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    if 'parity' in d: print('parity', d['parity'], d['ok'])
    elif d['generic']==0: print(d['config'], 'path', round(d['path_ms']*1000,1), 'json', round(d['json_ms']*1000,1), 'GB/s', round(d['gbps']), 'generic_tiles', d.get('generic_tiles'))
"
done
