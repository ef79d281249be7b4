python -m pytest tests/test_loader_gpu.py -x -q 2>&1 | tail -15
TEF_BENCH_TRAIN=0 python bench.py --steps 10 --warmup 3 > gpurun_out/bench41.log 2>&1; tail -c 600 gpurun_out/bench41.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench41.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_packed','loss')})
PY
TEF_BENCH_TRAIN=0 python bench.py --steps 10 --warmup 3 --workload iterative_128x128_b8_f4 > gpurun_out/bench41b.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench41b.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_packed','loss')})
else: print(open('gpurun_out/bench41b.log').read()[-1500:])
PY
