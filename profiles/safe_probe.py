"""Cost of the row-shifted softmax form in the table-path attention kernel (K1g): stage times with the form forced."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib, ops
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
    t16 = ops.pack_rows_f16(table)

def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

with torch.no_grad():
    for safe in (-1, 0, 1, -1):
        lib.nrms_set_option(b"attn_safe_softmax", safe)
        tu = timed(lambda: model.user_encoder.forward_indexed(t16, hist))
        tn = timed(lambda: model.get_news_vector({"title": tokens}))
        print(f"attn_safe_softmax={safe:2d}: users {tu:.3f} ms   news {tn:.3f} ms")
    lib.nrms_set_option(b"attn_safe_softmax", -1)
