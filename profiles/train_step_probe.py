"""One training step (B=128, 1+4 candidates, 50 history, dropout 0.2) for the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic
from newsrecommendationsystem_b200.train import TrainStep
dev = "cuda"
sd = synthetic.init_state_dict(seed=0)
m = NRMS(NRMSConfig); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.to(dev).train()
ts = TrainStep(m, lr=1e-4)
news = synthetic.make_news(65238)
cand, clicked = synthetic.make_train_batch(128, news, k_neg=4)
titles = torch.from_numpy(np.concatenate([cand, clicked], axis=1)).pin_memory()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(n):
    loss = ts.step_tokens(titles, 5)
torch.cuda.synchronize()
print("loss", float(loss))
