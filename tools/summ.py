import sys, json
tag = sys.argv[1]
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print(tag, 'ms/step', round(d['ms_per_step'], 4), 'EvI ms', round(d['roofline']['ms_per_launch'], 4), 'frac', round(d['roofline']['frac'], 3), 'value', round(d['value'], 1), flush=True)
    elif 'rror' in l or 'assert' in l.lower():
        print(tag, l.strip())
