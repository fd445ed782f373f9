for boot in 1 0; do
echo "== BOOT=$boot"
VQ_MMA_BOOT=$boot timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('main B=%d value=%.0f ms=%.4f e2e=%.0f kernel_ms=%.4f frac=%.3f unc=%s launches=%s' % (d['config']['batch'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['uncertified_queries_per_batch'], d['gpu_launches']))
        for m in d['sweep']: print('  sweep B=%d value=%.0f ms=%.4f kernel_ms=%.4f %s frac=%.3f unc=%s' % (m['batch'], m['value'], m['ms_per_step'], m['roofline']['kernel_ms'], m['roofline']['bound'], m['roofline']['frac'], m['uncertified_queries_per_batch']))
    else: print(l.rstrip()[:300])
"
done
