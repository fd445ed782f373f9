for pdl in 0 8 2 10 16 1 31; do
echo "== PDL mask=$pdl"
VQ_PDL=$pdl timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-hnsw --batch 32 --no-sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('main B=%d value=%.0f ms=%.4f e2e=%.0f kernel_ms=%.4f' % (d['config']['batch'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']))
"
done
