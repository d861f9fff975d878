"""A few calls of the single-user path for the ncu launch list (per-kernel durations of head_kernel / pool_score_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = NRMS(NRMSConfig).to(dev).eval().set_precision("fp32")
n_rows = 65239
table = torch.randn(n_rows, 300, device=dev) * 0.4
rng = np.random.default_rng(0)
hist = torch.from_numpy(rng.integers(0, n_rows, 50).astype(np.int32)).to(dev)
ue = m.user_encoder
w = (*ue.multihead_self_attention.packed(), ue.additive_attention.linear.weight, ue.additive_attention.linear.bias,
     ue.additive_attention.attention_query_vector)
for C in (37, 300, 4096):
    cand = torch.from_numpy(rng.integers(0, n_rows, C).astype(np.int32)).to(dev)
    for _ in range(3):
        ops.recommend_user(table, hist, cand, *w)
torch.cuda.synchronize()
