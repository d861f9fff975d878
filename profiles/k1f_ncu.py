"""One user-encoder call (and one news-encoder call) at the bench's full size: the command ncu profiles."""
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
with torch.no_grad():
    table = torch.zeros((news.shape[0] + 1, 300), device=dev)
    for _ in range(2):
        table[:-1] = model.get_news_vector({"title": tokens})
        uv = model.user_encoder.forward_indexed(table, hist)
torch.cuda.synchronize()
print("ok", float(uv.abs().sum()))
