for rep in 1 2 3; do for wl in iterative_240x320_1Mev iterative_128x128_b8_f4 iterative_480x640_1Mev; do
  TEF_BENCH_TRAIN=0 python bench.py --workload $wl --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
x=json.loads(sys.stdin.read()); print(x['config']['workload'], round(x['ms_per_step'],3), round(x['value'],1), 'e2e', round(x['e2e']['value'],1), round(x['e2e']['ms_per_step'],3), 'packed', round(x['e2e_packed']['value'],1), round(x['e2e_packed']['ms_per_step'],3))"
done; done
