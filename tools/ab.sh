for v in "" registrar_b200/ab/t128.so registrar_b200/ab/t128b.so registrar_b200/ab/t512.so registrar_b200/ab/p6.so; do
  echo "== ${v:-base}"
  REGK_LIB=$v timeout 300 python tools/quick_time.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    if d['generic']==0: print(d['config'], round(d['path_ms']*1000,1), round(d['json_ms']*1000,1), round(d['gbps']))
"
done
