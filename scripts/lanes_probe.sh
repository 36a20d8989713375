python -m pytest tests/test_gpu_ospline.py -x -q -k "lanes" 2>&1 | tail -15
for L in ${LANES_LIST:-1 2 3 4 6 8}; do
  python bench.py --steps 5 --warmup 3 --no-predict --no-cpu --no-dense --no-fit --no-grad --lanes $L > gpurun_out/osp_lanes_$L.log 2>&1
  python -c "
import json
for l in open('gpurun_out/osp_lanes_$L.log'):
    if l.startswith('{'):
        d=json.loads(l); print('lanes $L', round(d['value']), round(d['e2e']['value']), d['config']['newton_iters_per_eval'], d['sharded_vs_single_gpu']['max_rel_logpost'], d['batch_entry_point']['value'])
"
done
