#!/bin/bash
# bench line + per-launch profile of the product library and of every experimental build (opensbli_b200/libosbli_b200_<tag>.so)
T=${1:-r2v}
mkdir -p gpurun_out
for lib in opensbli_b200/libosbli_b200.so opensbli_b200/libosbli_b200_*.so; do
  [ -e "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libosbli_b200//; s/^_//'); tag=${tag:-product}
  OSB_PROFILE_LIST=1 OSB_B200_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 5 > gpurun_out/${T}_${tag}.json 2> gpurun_out/${T}_${tag}.err
  python - <<PY
import json
d = [json.loads(l) for l in open('gpurun_out/${T}_${tag}.json') if l.startswith('{')][-1]
print('$tag', 'ms/step %.2f' % d['ms_per_step'], {k: round(v, 2) for k, v in d['roofline']['families_ms'].items() if v}, (d.get('parity') or {}).get('max_rel_err'))
PY
  grep -h "osb_profile" gpurun_out/${T}_${tag}.err | head -3 | cut -c1-400
done
