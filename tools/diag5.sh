timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -2
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -m gpu 2>&1 | tail -3
VQ_FINISH_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep --no-graph 2>&1 | grep "finish dbg" | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('main B=%d value=%.0f ms=%.4f e2e=%.0f kernel_ms=%.4f frac=%.3f unc=%s launches=%s' % (d['config']['batch'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['uncertified_queries_per_batch'], d['gpu_launches']))
        for m in d['sweep']: print('  sweep B=%d value=%.0f ms=%.4f kernel_ms=%.4f %s frac=%.3f unc=%s' % (m['batch'], m['value'], m['ms_per_step'], m['roofline']['kernel_ms'], m['roofline']['bound'], m['roofline']['frac'], m['uncertified_queries_per_batch']))
    else: print(l.rstrip()[:300])
"
