for r in 1 2; do for hc in ${HCS:-0 128 256 384}; do
R3D_BENCH_OPTIONS="host_chunk=$hc" python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/hc_$hc.json 2> gpurun_out/hc_$hc.err
python - <<PY
import json
d=json.load(open('gpurun_out/hc_$hc.json')); e=d['e2e']
print('host_chunk=$hc', 'value', round(d['value']), 'e2e', round(e['value']), 'sync', round(e['sync_value']), round(e['sync_ms_per_step'],3))
PY
done; done
