"""A/B of the user-encoder softmax form on the table path: -1 = bound pass (qk_bound_kernel) choosing 2^s or the
row-shifted form per call, 1 = always row-shifted (no bound pass).  Whole evaluate pass, bench N = 1 size, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for opt in (-1, 1, -1, 1):
    lib.nrms_set_option(b"attn_safe_softmax", opt)
    ts = []
    for i in range(8):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); means = evaluate_tensors(model, inputs); b.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    print(f"attn_safe_softmax {opt:2d}: {np.mean(ts):.3f} ms per evaluate pass  (auc {means[0]:.6f})", flush=True)
