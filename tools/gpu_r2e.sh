#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_bank_ring.py tests/test_gpu_gemm.py tests/test_gpu_x3.py tests/test_gpu_install.py -q -x -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn" | tail -60 > $out/r2f_new.log; tail -25 $out/r2f_new.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > $out/r2f_gpu_tests.log
tail -8 $out/r2f_gpu_tests.log
python tools/trace_step.py --out $out/r2f_trace_n1.txt > /dev/null 2>&1 || echo trace failed
python bench.py --steps 20 > $out/r2f_bench_n1.json 2> $out/r2f_bench_n1.err || tail -20 $out/r2f_bench_n1.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2f_bench_n1.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',l['e2e']['value'], l['e2e']['serial_steps_per_s'], l['e2e']['eager_module_api_steps_per_s']); r=l['roofline']; print('roofline',r['achieved'], r['frac'], r['avg_launch_ms']); print(l['modes']); print('eval',l['eval']['resident_ms'],l['eval']['e2e_ms'])
PY
