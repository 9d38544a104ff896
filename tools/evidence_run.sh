#!/bin/bash
# round-2 evidence run (one GPU): tests, bench both arms, ncu launch list + full captures -> gpurun_out/r02_*
# (then: python tools/summarize_profiles.py r02 ... ; python tools/sass_report.py r02 ; python tools/fill_docs.py r02)
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/r02_pytest_gpu.log; cat $O/r02_pytest_gpu.log
KVG_CLOCKS_CSV=$O/r02_clocks_n1.csv timeout 900 python bench.py --steps 20 --warmup 5 2>$O/r02_bench_n1.err > $O/r02_bench_n1.json; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>$O/r02_bench_ref.err > $O/r02_bench_reference_arm.json; echo "ref exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'step_ms', round(d['ms_per_step'], 4), 'value', d['value'], 'launches/step', d['gpu_launches'] / d['steps'], 'parity', d['parity']['status'])
print('kernels_us', {k: round(v * 1e3, 1) for k, v in d['kernel_ms_per_step'].items()})
print('roofline', json.dumps(d['roofline'])[:600])
print('big', json.dumps(d['roofline_hbm_bound'])[:2500])
print('e2e', json.dumps(d['e2e'])[:500]); print('other', json.dumps(d['other_configs'])[:900]); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
r = json.loads(open('gpurun_out/r02_bench_reference_arm.json').read().strip().splitlines()[-1]); print('ref', r.get('value'), r.get('cpu_baseline'))
PY
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $B --big-files 0 --big-records 0 > $O/ncu_b.log 2>&1; tail -1 $O/ncu_b.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_order|k_classify|k_pciids' -s 33 -c 12 -o $O/r02_cfg2 -f $B --big-files 0 --big-records 0 > $O/ncu_c.log 2>&1; tail -1 $O/ncu_c.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pciids -s 9 -c 3 -o $O/r02_parse256 -f python tools/time_parse.py > $O/ncu_p.log 2>&1; tail -1 $O/ncu_p.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_classify_ragged|k_pack_survivors|k_order_scatter|k_order_hist|k_order_heads' -c 12 -o $O/r02_big -f python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --big-files 0 --records 16 > $O/ncu_g.log 2>&1; tail -1 $O/ncu_g.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_classify_send|k_shard_gather' -s 12 -c 2 -o $O/r02_shard -f python tools/time_shard.py > $O/ncu_s.log 2>&1; tail -1 $O/ncu_s.log | cut -c1-200
ls -la $O/r02_*
