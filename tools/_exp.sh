for m in 0 1 2 3; do DSVG_EXP=$m DSVG_NO_CLOCKS=1 DSVG_BENCH_TIMEOUT=200 timeout 220 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('exp$m', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()});
[print('   ', s['MNK'], s['launches'], s['ms'], s['tflops']) for s in d['roofline']['top_shapes']]"; done
