"""Debug: per-phase clock64 trace of K1 v2 (worker thread 0 of block 0)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib, ops
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = "cuda"
sd = synthetic.init_state_dict(num_words=5001, seed=0)
class Cfg(NRMSConfig): num_words = 5001
m = NRMS(Cfg); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.to(dev).eval()
lib = _lib.load()
buf = (ctypes.c_longlong * 2048)()
with torch.no_grad():
    if S == 50:
        x = torch.randn(1184, 50, 300, device=dev) * 0.3
        for _ in range(2): m.get_user_vector(x)
    else:
        toks = torch.from_numpy(synthetic.make_news(2960, num_words=5001)).to(dev)
        for _ in range(2): m.get_news_vector({"title": toks})
    torch.cuda.synchronize()
    lib.nrms_debug_read_trace(buf, 2048)
    if S == 50: m.get_user_vector(x)
    else: m.get_news_vector({"title": toks})
    torch.cuda.synchronize()
n = lib.nrms_debug_read_trace(buf, 2048)
ev = [(buf[i] >> 48, buf[i] & 0xFFFFFFFFFFFF) for i in range(n)]
names = {1: "W1 start", 2: "qk arrived", 3: "s_ready", 4: "p arrived", 5: "o_ready", 6: "tile start", 7: "gather done", 8: "acc_full"}
t0 = ev[0][1]
prev = t0
for tag, t in ev[:90]:
    print(f"{names.get(tag, tag):12s} t={t - t0:8d} d={t - prev:6d}")
    prev = t
# per-phase averages
import collections
d = collections.defaultdict(list)
for (a, ta), (b, tb) in zip(ev[:-1], ev[1:]):
    d[(a, b)].append(tb - ta)
for k, v in sorted(d.items()):
    print(names.get(k[0]), "->", names.get(k[1]), "n", len(v), "mean", int(np.mean(v)), "min", min(v), "max", max(v))
