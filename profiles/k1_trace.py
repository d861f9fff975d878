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
buf = (ctypes.c_longlong * 2048)()  # build with NRMS_K1_TRACE=1 python newsrecommendationsystem_b200/csrc/build.py --force
with torch.no_grad():
    if S == 50:
        x = torch.randn(1184, 50, 300, device=dev) * 0.3
        for _ in range(2): m.get_user_vector(x)
    else:
        toks = torch.from_numpy(synthetic.make_news(2960, num_words=5001)).to(dev)
        for _ in range(2): m.get_news_vector({"title": toks})
    torch.cuda.synchronize()
    if S == 50: m.get_user_vector(x)
    else: m.get_news_vector({"title": toks})
    torch.cuda.synchronize()
fn = getattr(lib, 'nrms_debug_read_trace3', None) or lib.nrms_debug_read_trace
n = fn(buf, 2048)
ev = [(buf[i] >> 48, buf[i] & 0xFFFFFFFFFFFF) for i in list(range(n)) + list(range(250, 250 + n))]
names = {20: "pass start", 21: "acc_full", 22: "W1a done", 23: "W1b done", 24: "s_ready a", 25: "s_ready b", 26: "W2a done", 27: "W2b done", 28: "o_ready a", 29: "o_ready b", 30: "W3a done", 31: "W3b done", 11: "ld done", 12: "stores done", 13: "inv", 14: "O ld done", 15: "O stored", 1: "W1 start", 2: "qk arrived", 3: "s_ready", 4: "p arrived", 5: "o_ready", 6: "tile start", 7: "gather done", 8: "acc_full"}
import collections
for role in (0, 1):
    evr = sorted([(tag - 100 * role, t) for tag, t in ev if (tag >= 100) == bool(role)], key=lambda x: x[1])
    print("=== role", role, "events", len(evr))
    d = collections.defaultdict(list)
    for (a, ta), (b, tb) in zip(evr[:-1], evr[1:]):
        d[(a, b)].append(tb - ta)
    for k, v in sorted(d.items()):
        print(names.get(k[0], k[0]), "->", names.get(k[1], k[1]), "n", len(v), "mean", int(np.mean(v)), "min", min(v), "max", max(v))
