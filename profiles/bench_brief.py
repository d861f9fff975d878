"""Print the headline numbers of bench.py JSON lines read from stdin (or files given as arguments)."""
import json, sys
src = sys.stdin if len(sys.argv) < 2 else (l for f in sys.argv[1:] for l in open(f))
for l in src:
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    if d.get("impl") == "reference":
        print("reference:", d.get("value"), d.get("unit"), d.get("cpu_baseline", {}).get("sample", "")[:80])
        continue
    print(f"n_gpus {d['n_gpus']}  ms/step {d['ms_per_step']:.3f}  value {d['value']:.4g} {d['unit']}  e2e {d['e2e']['value']:.4g} ({d['e2e'].get('ms_per_step', 0):.3f} ms)")
    print("  stage_ms", {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()})
    for key in ("roofline", "roofline_news"):
        r = d.get(key)
        if r:
            print(f"  {key}: {r['kernel'][:60]}  {r.get('us_per_launch', 0):.1f} us  frac {r['frac']:.3f} ({r['achieved']:.0f} {r['unit']})")
    if d.get("train"):
        print("  train", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["train"].items() if k in ("samples_per_s", "ms_per_step", "loss")})
    for k in ("mind_large", "train_ln", "cpu_baseline"):
        if d.get(k):
            print(f"  {k}:", json.dumps(d[k])[:300])
