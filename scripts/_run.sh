set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -c 300 gpurun_out/bench_r1j.err
for wl in iterative_480x640_100kev iterative_480x640_250kev iterative_480x640_500kev iterative_480x640_2Mev iterative_480x640_4Mev iterative_480x640_1Mev_edges iterative_240x320_1Mev iterative_128x128_b8_f1 iterative_128x128_b8_f4 linear_128x128_b8_f4 linear_480x640_1Mev; do
  TEF_BENCH_TRAIN=0 python bench.py --workload $wl --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1j_$wl.json
done
for wl in inference_480x640_1Mev validation_480x640_100kev validation_480x640_500kev train_128x128_b8; do python bench.py --workload $wl --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1j_$wl.json; done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1j_reference.json
ncu --set full --clock-control none --import-source on -k regex:"iter_|sort_scatter|update_pass" -s 12 -c 4 -o gpurun_out/prof_r1_j -f python scripts/profile_step.py --workload iterative_480x640_1Mev --steps 3 > gpurun_out/ncu_full_j.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_j.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list_j.log 2>&1
ls -la gpurun_out | tail -4
