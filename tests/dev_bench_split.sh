python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('default: value %.0f ms %.4f whole %.3f scan %.3f' % (d['value'], d['ms_per_step'], d['roofline']['whole_step_frac'], d['roofline']['frac']), d['phases_ms'], 'e2e', d['e2e']['value'])
"
