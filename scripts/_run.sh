python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -c 300 gpurun_out/bench_r1i.err
