"""Evaluate passes over MIND-small-shaped synthetic data for the ncu launch list / --set full captures
(same model, data and call as bench.py's timed region).  Usage: python profiles/eval_step_probe.py [n_passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic
from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
dev = torch.device("cuda", 0)
sd = synthetic.init_state_dict(num_words=bench.NUM_WORDS, seed=0)
model = NRMS(NRMSConfig)
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
model.to(dev).eval().set_precision("tf32")
news, imp = bench.make_data(1)
host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
inputs = EvalInputs.from_host(host, dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    means = evaluate_tensors(model, inputs)
torch.cuda.synchronize()
print("metrics", means)
