"""One-line summary of a bench.py JSON line on stdin (value, ms per step, per-class ms, extend / shadow rays per second, e2e)."""
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; c=r["rank0_class_ms_per_step"]
        print(f"value={d['value']:.0f} Mrays/s  ms/step={d['ms_per_step']:.2f}  extend={c['extend']:.2f} shade={c['shade']:.2f} shadow={c['shadow']:.2f} other={c['other']:.2f}  ext={r['rays_per_s']/1e9:.2f}G/s sh={r['shadow']['rays_per_s']/1e9:.2f}G/s frac={r['frac']:.3f} e2e={d['e2e']['value']:.0f}")
    else: print(l.rstrip())
