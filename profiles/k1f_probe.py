"""Timing probe of the fused table-path kernel (K1f): user / news encoder stage times at the bench's full size for a list
of `k1f_debug` component-removal masks (results are garbage for mask != 0).  PROBE_MASKS="0 1 2 ..." PROBE_OLD=1 adds
the two-kernel path (K1g + K2)."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
sd = synthetic.init_state_dict(num_words=bench.NUM_WORDS, seed=0)
model = NRMS(NRMSConfig); model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); model.to(dev).eval().set_precision("tf32")
news, imp = bench.make_data(1)
tokens = torch.from_numpy(news).to(dev)
hist = imp["hist_rows"].copy(); hist[hist < 0] = news.shape[0]
hist = torch.from_numpy(hist.astype(np.int32)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    table = torch.zeros((news.shape[0] + 1, 300), device=dev)
    table[:-1] = model.get_news_vector({"title": tokens})

def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

masks = [int(x) for x in os.environ.get("PROBE_MASKS", "0 1 2 4 8 16 32 3 63").split()]
with torch.no_grad():
    if os.environ.get("PROBE_OLD", "1") != "0":
        lib.nrms_set_option(b"fused_pool", 0)
        print(f"K1g+K2   users {timed(lambda: model.user_encoder.forward_indexed(table, hist)):.3f} ms   news {timed(lambda: model.get_news_vector({'title': tokens})):.3f} ms")
        lib.nrms_set_option(b"fused_pool", 1)
    for safe in (-1, 1):
        lib.nrms_set_option(b"attn_safe_softmax", safe)
        for m in masks:
            lib.nrms_set_option(b"k1f_debug", m)
            lib.nrms_set_option(b"time_k1", 1)
            tu = timed(lambda: model.user_encoder.forward_indexed(table, hist), reps=3)
            ku = lib.nrms_get_stat(b"k1g_ms") / max(1.0, lib.nrms_get_stat(b"k1g_launches"))
            lib.nrms_set_option(b"time_k1", 1)
            tn = timed(lambda: model.get_news_vector({"title": tokens}), reps=3)
            kn = lib.nrms_get_stat(b"k1gn_ms") / max(1.0, lib.nrms_get_stat(b"k1gn_launches"))
            lib.nrms_set_option(b"time_k1", 0)
            print(f"K1f safe={safe:2d} dbg={m:2d}: users {tu:.3f} ms (kernel {ku:.3f})   news {tn:.3f} ms (kernel {kn:.3f})")
        if os.environ.get("PROBE_SAFE", "0") == "0":
            break
    lib.nrms_set_option(b"k1f_debug", 0)
    lib.nrms_set_option(b"attn_safe_softmax", -1)
