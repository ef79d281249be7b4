#!/bin/bash
# round 2, GPU call M: full GPU test-suite, smoke, default bench line, launch list (ncu) of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2m_pytest.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m.err; tail -c 1200 gpurun_out/r2m_bench.json; tail -5 gpurun_out/r2m.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>> gpurun_out/r2m.err; cut -c1-400 gpurun_out/r2m_bench_reference.json
TEF_BENCH_TRAIN=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2m_ncu_bench.log 2>&1
tail -3 gpurun_out/r2m_launches.csv | cut -c1-200
