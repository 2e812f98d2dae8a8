#!/bin/bash
# A/B of one experiment environment variable (experiments build): scripts/ab_env.sh VAR "v1 v2 ..." [repeats]
# prints value, ms/step and the per-launch times of the serialised profiling pass for every setting.
var=$1; vals=$2; reps=${3:-2}
mkdir -p gpurun_out
for r in $(seq $reps); do for v in $vals; do
  env $var=$v python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/ab_${var}_$v.json 2> gpurun_out/ab_${var}_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/ab_${var}_$v.json'))
print('$var=$v', round(d['value']), round(d['ms_per_step'], 4), ' '.join('%s=%.1f' % (L['name'][:24], L['ms'] * 1e3) for L in d['roofline']['launches']))
PY
done; done
