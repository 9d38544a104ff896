#!/bin/bash
# round 2, GPU call 4: K1 (32-byte cells, direct table) + new orderings (kvg_order.cuh): parity, bench, ncu
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 --big-files 256 2>gpurun_out/run4_bench.err > gpurun_out/run4_bench.json
python - <<PY
import json
d = json.loads(open('gpurun_out/run4_bench.json').read().strip().splitlines()[-1])
p = d['roofline_hbm_bound']['pciids_parse']
print('step_ms', round(d['ms_per_step'], 4), 'parse256_GBps', round(p['achieved'], 1), 'frac', round(p['frac'], 3),
      'kernels_us', {k: round(v * 1e3, 1) for k, v in d['kernel_ms_per_step'].items()})
print('e2e', d['e2e']['ms_per_step'], 'scan16M', d['roofline_hbm_bound']['classify_compact']['whole_scan_ms'], 'launches/step', d['gpu_launches']/d['steps'])
PY
tail -3 gpurun_out/run4_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pciids_scan -s 1 -c 1 -o gpurun_out/r02d_parse_k1 python tools/profile_kernels.py 65536 256 2 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3 --big-files 0 --big-records 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_order|k_classify|k_pack|k_tile' -s 30 -c 30 --csv --log-file gpurun_out/r02d_launches16m.csv python tools/profile_kernels.py 16777216 1 4 > /dev/null 2>&1
tail -3 gpurun_out/r02d_launches16m.csv
