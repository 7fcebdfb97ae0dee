# After `gpurun -- bash scripts/round_artifacts.sh`: turn gpurun_out/r2g_* into the committed profiles/ entries (run here).
set -e
cd "$(dirname "$0")/.."
ncu -i gpurun_out/r2g_step_full.ncu-rep --page raw --csv > profiles/r2_step_raw.csv 2>/dev/null
python scripts/ncu_traffic.py profiles/r2_step_raw.csv --config 2 --regime clustered --launches-per-step 8 \
    --note "ncu --set full --clock-control none --import-source on -s 8 -c 8 python scripts/one_step.py (second pass of the stage), round 2 final build"
cp gpurun_out/r2g_launches_config2.csv profiles/r2_launches_config2.csv
cp gpurun_out/r2g_bench_c2.json profiles/r2_bench_config2.json
{
  echo '`ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python bench.py --steps 3 --warmup 3 --no-e2e --no-pipelined --no-cpu-baseline --no-graph` (round-2 final build; the same command exited 0 without ncu immediately before).  ncu serialises the launches and runs them cold, so the SHARES are what to read; bench.py gives the timed numbers.'
  echo; echo '```'; python scripts/launch_summary.py profiles/r2_launches_config2.csv 16; echo '```'
} > profiles/r2_launches_config2.md
SO=maskrcnn_tf2_b200/libmrcnn_roi_b200.so
for pair in "nms_sweep_kernelILi3ELb1:r2_sass_nms_sweep_3_unit" "topk_cluster_kernelILi2:r2_sass_topk_cluster_2" "roialign_fwd_kernelILi2ELi1:r2_sass_roialign_fwd_2_1"; do
  pat=${pair%%:*}; out=${pair##*:}
  fn=$(cuobjdump -sass $SO 2>/dev/null | grep "Function :" | sed 's/.*Function : //' | grep "$pat" | head -1)
  cuobjdump -sass -fun "$fn" $SO 2>/dev/null | grep -v "^cuobjdump warning" > profiles/$out.txt
done
python - <<'PY'
import json, bench
l=[x for x in open('gpurun_out/r2g_bench_c2.json') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']})
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'],
      'pipelined', d.get('pipelined',{}).get('value'), 'eager', d['eager']['value'])
print('traffic hash ok:', bench.source_hash() == json.load(open('profiles/roofline_traffic.json'))['source_sha256'])
PY
tail -9 profiles/r2_launches_config2.md | head -8
tail -2 gpurun_out/r2g_tests.log; grep "^final" gpurun_out/r2g_layers.txt
