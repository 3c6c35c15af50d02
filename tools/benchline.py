import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(round(d['value'] / 1e6, 1), 'Mpairs/s', round(d['ms_per_step'], 4), 'ms/step  frac', round(d['roofline']['frac'], 4), {k: round(v, 4) for k, v in d['roofline']['kernel_ms'].items()})
