"""Print the headline + sweep of one or more bench JSON lines side by side."""
import json, sys
ds = [json.loads(open(f).read().strip().splitlines()[-1]) for f in sys.argv[1:]]
for f, d in zip(sys.argv[1:], ds):
    print(f"{f}: {d['value']/1e6:.2f} M maps/s  step {d['ms_per_step']*1e3:.2f} us  bwd {d['roofline']['frac']:.3f} fwd {d['roofline_forward']['frac']:.3f} step {d['roofline_step']['frac']:.3f}")
n = len(ds[0].get('sweep', []))
for i in range(n):
    row = ds[0]['sweep'][i]['workload'][22:].replace(' pad=1 reflect', '').replace(' pad=2 reflect', '')
    print(f"{row:28s}", " | ".join(f"f {d['sweep'][i]['us_fwd']:6.1f} b {d['sweep'][i]['us_bwd']:6.1f} {d['sweep'][i]['step_frac']:.3f}" for d in ds))
