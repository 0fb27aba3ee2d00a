#!/bin/bash
# Turn the captures tools/profile_r2.sh left in gpurun_out/ into the tracked summaries under profiles/ (no GPU needed).
set -e
cp gpurun_out/r2_launches.csv profiles/r2_launches.csv
cp gpurun_out/r2_sanitizer.txt profiles/r2_sanitizer.txt
python tools/ncu_traffic.py config3:10000000=gpurun_out/prof_r2_config3.ncu-rep config2:1000000=gpurun_out/prof_r2_config2.ncu-rep \
    config5:12500000=gpurun_out/prof_r2_config5.ncu-rep
for c in config3 config2 config5; do
  python tools/ncu_summary.py gpurun_out/prof_r2_$c.ncu-rep 14 > profiles/r2_ncu_$c.txt
done
{
  echo "# SASS opcode histogram of the two kernels (executed warp instructions, ncu --set full capture of config 3, 10 M records;"
  echo "# tools/sassmix.py).  UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier, ATOMS/ATOMG/RED = the totals."
  for k in regk_path_kernel regk_json_kernel; do
    echo "== $k"
    python tools/sassmix.py gpurun_out/prof_r2_config3.ncu-rep $k 312500
  done
  echo "== static: TMA / mbarrier / byte-masked bulk-store opcodes in libregk.so (cuobjdump -sass)"
  cuobjdump -sass registrar_b200/libregk.so | grep -o -E "UBLKCP[A-Z_.]*|SYNCS[A-Z_.]*|UTMA[A-Z]*" | sort | uniq -c
} > profiles/r2_sass_opcodes.txt
ls -la profiles/r2_*
