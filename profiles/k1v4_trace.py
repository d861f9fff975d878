"""Debug: clock64 phase trace of K1 v5 (NRMS_K1_VARIANT=4: v4), block 0 (build with NRMS_K1_TRACE=1 python .../csrc/build.py --force).
Tracers: 0 worker role 0 (warp 2), 1 worker role 1 (warp 6), 2 projection MMA issuer, 3 attention MMA issuer."""
import ctypes, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = "cuda"
sd = synthetic.init_state_dict(num_words=5001, seed=0)
class Cfg(NRMSConfig): num_words = 5001
m = NRMS(Cfg); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.to(dev).eval()
lib = _lib.load()
with torch.no_grad():
    if S == 50:
        x = torch.randn(2368, 50, 300, device=dev) * 0.3
        f = lambda: m.get_user_vector(x)
    else:
        toks = torch.from_numpy(synthetic.make_news(5920, num_words=5001)).to(dev)
        f = lambda: m.get_news_vector({"title": toks})
    for _ in range(3): f()
    torch.cuda.synchronize()
NT = 5 if os.environ.get('NRMS_K1_VARIANT', '6') in ('5', '6') else 4
buf = (ctypes.c_longlong * (1024 * NT))(); cnt = (ctypes.c_int * NT)()
{'6': lib.nrms_debug_read_trace6, '5': lib.nrms_debug_read_trace5}.get(os.environ.get('NRMS_K1_VARIANT', '6'), lib.nrms_debug_read_trace4)(buf, cnt)
names = {20: "pass start", 21: "acc_full", 22: "W1 done", 23: "W3fin done", 24: "s_ready a", 25: "s_ready b", 26: "W2a done",
         27: "W2b done", 28: "o_ready a", 29: "o_ready b", 30: "W3 done", 31: "staged",
         40: "P pass start", 41: "P acc_empty", 42: "P kc0", 43: "P kc1", 44: "P kc2", 45: "P kc3", 46: "P kc4",
         47: "P wait", 48: "P issued", 55: "A S issued", 56: "A O issued", 60: "G free0", 61: "G free1", 62: "G free2", 63: "G free3", 64: "G free4", 70: "G full0", 71: "G full1", 72: "G full2", 73: "G full3", 74: "G full4", 50: "A pass start", 51: "A kv a", 52: "A kv b", 53: "A p a", 54: "A p b"}
allev = []
for who in range(NT):
    ev = [(buf[who * 1024 + i] >> 48, buf[who * 1024 + i] & 0xFFFFFFFFFFFF) for i in range(cnt[who])]
    allev += [(t, who, tag) for tag, t in ev]
    print("=== tracer", who, "events", len(ev))
    d = collections.defaultdict(list)
    for (a, ta), (b, tb) in zip(ev[:-1], ev[1:]):
        d[(a, b)].append(tb - ta)
    for k, v in sorted(d.items()):
        print(f"  {names.get(k[0], k[0]):>14s} -> {names.get(k[1], k[1]):<14s} n {len(v):4d} mean {int(np.mean(v)):6d} min {min(v):6d} max {max(v):6d}")
# merged timeline of the 3rd tile (steady state)
allev.sort()
t0 = allev[0][0]
starts = [t for t, who, tag in allev if who == 0 and tag == 20]
if len(starts) > 24:
    lo, hi = starts[14], starts[17]
    print("=== merged timeline, passes 14..16 (cycles since pass-14 start; pass 16 opens a tile)")
    for t, who, tag in allev:
        if lo <= t <= hi:
            print(f"  {t - lo:7d}  tracer {who}  {names.get(tag, tag)}")
