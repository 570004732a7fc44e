for c in 1.5 1.2 1.0 0.8; do
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --legs dense_scene --sequences '' --map-cell $c > gpurun_out/cell_$c.json 2> gpurun_out/cell_$c.err
  python - <<P
import json
l=json.loads([x for x in open('gpurun_out/cell_$c.json') if x.startswith('{')][-1])
print('$c', round(l['value']), round(l['value_l2_warm']), round(l['full_scan']['value']), l['roofline']['single_pass_kernel'], round(l['dense_scene']['value']), round(l['e2e']['value']))
P
done
