import sys, json
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
    except Exception:
        print(line.rstrip()); continue
    extra = {k: (('%.2e' % v) if isinstance(v, float) else v) for k, v in d.items() if k not in ('kernel', 'variant', 'us') and v is not None}
    print('%-12s %-62s %8.1f us  %s' % (d['kernel'], d['variant'], d.get('us', -1), extra))
