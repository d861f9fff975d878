"""Two evaluate passes at the bench's N = 1 size (the command ncu profiles; the first pass is the warm-up)."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib
from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
lib = _lib.load()
dev = torch.device("cuda", 0)
sd = synthetic.init_state_dict(num_words=bench.NUM_WORDS, seed=0)
model = NRMS(NRMSConfig); model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); model.to(dev).eval().set_precision("tf32")
news, imp = bench.make_data(1)
host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
inputs = EvalInputs.from_host(host, dev)
for _ in range(2):
    means = evaluate_tensors(model, inputs)
torch.cuda.synchronize()
print("means", means)
