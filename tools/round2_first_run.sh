#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root, one GPU):
#   gpurun --timeout 900 -- 'bash tools/round2_first_run.sh > gpurun_out/round2_first.log 2>&1'
# 1. the suites that were written after round 1's GPU budget ran out (never executed on a GPU):
#      tests/test_serve_gpu.py            Register -> ListAndWatch -> Allocate on the real scan
#      KVG_PARSE=v2 parse parity          csrc/kvg_parse_v2.cuh, the barrier-free pci.ids parse
# 2. A/B of the two parse kernels on the HBM-bound leg (128 images) and on the config-2 step.
set -x
python -m pytest tests/test_serve_gpu.py -m gpu -x -q 2>&1 | tail -5
KVG_PARSE=v2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "kats or golden or general or fuzz or tile or batch or million or mdev or discovery or native" 2>&1 | tail -5
for v in v1 v2; do
  KVG_PARSE=$v python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = d['roofline_hbm_bound']['pciids_parse']
print('$v', 'step_ms', round(d['ms_per_step'], 4), 'parse128_GBps', round(p['achieved'], 1), 'frac', round(p['frac'], 3),
      'kernels_us', {k: round(v * 1e3, 1) for k, v in d['kernel_ms_per_step'].items() if k.startswith('pciids')})"
done
# 3. the warp-per-digit tile scan (csrc/kvg_radix_exp.cuh): parity, then the config-2 step
KVG_TILESCAN=warp python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_pci or orderings or million or mdev or properties" 2>&1 | tail -3
for v in block warp; do
  KVG_TILESCAN=$v python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tilescan=$v', 'step_ms', round(d['ms_per_step'], 4), 'radix_tilescan_us', round(d['kernel_ms_per_step']['radix_tilescan'] * 1e3, 1),
      'scan16M_ms', round(d['roofline_hbm_bound']['classify_compact']['whole_scan_ms'], 4))"
done
# 4. the contiguous-ownership scatter (csrc/kvg_radix_exp.cuh), alone and with the warp tile scan
KVG_SCATTER=c python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_pci or orderings or million or mdev or properties" 2>&1 | tail -3
for env in "KVG_SCATTER=c" "KVG_SCATTER=c KVG_TILESCAN=warp" "KVG_SCATTER=c KVG_TILESCAN=warp KVG_PARSE=v2"; do
  env $env python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$env', 'step_ms', round(d['ms_per_step'], 4), {k: round(v * 1e3, 1) for k, v in d['kernel_ms_per_step'].items() if k.startswith('radix') or k.startswith('pciids')})"
done
