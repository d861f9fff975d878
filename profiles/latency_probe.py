"""Latency of the single-user path: nrms_recommend_user (two cluster launches) against the library's
batched FP32 kernels at batch 1 and a torch-eager restatement of the reference's op sequence on the same GPU.
Timed with CUDA events over 200 back-to-back calls after warm-up (host launch overhead included, no sync inside)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, ops, _lib

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = NRMS(NRMSConfig).to(dev).eval().set_precision("fp32")
n_rows = 65239
table = torch.randn(n_rows, 300, device=dev) * 0.4
table[-1] = 0
rng = np.random.default_rng(0)
hist = torch.from_numpy(rng.integers(0, n_rows, 50).astype(np.int32)).to(dev)
ue = m.user_encoder
w = (*ue.multihead_self_attention.packed(), ue.additive_attention.linear.weight, ue.additive_attention.linear.bias,
     ue.additive_attention.attention_query_vector)
out = {}


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / n


def eager_reference(cand):
    """the reference's op sequence (user_encoder.py:15-26, multihead_self.py:15-76, additive.py:27-53, dot_product.py) in
    torch eager fp32 on the GPU"""
    x = table[hist.long()].unsqueeze(0)
    W = w[0]; b = w[1]
    q, k, v = (x @ W[i * 300:(i + 1) * 300].T + b[i * 300:(i + 1) * 300] for i in range(3))
    sh = lambda t: t.view(1, 50, 15, 20).transpose(1, 2)
    s = torch.exp(sh(q) @ sh(k).transpose(-1, -2) / np.sqrt(20))
    a = s / (s.sum(-1, keepdim=True) + 1e-8)
    c = (a @ sh(v)).transpose(1, 2).contiguous().view(1, 50, 300)
    t = torch.tanh(c @ w[2].T + w[3])
    al = torch.softmax(t @ w[4], dim=1)
    u = torch.bmm(al.unsqueeze(1), c).squeeze(1)
    sc = torch.bmm(table[cand.long()].unsqueeze(0), u.unsqueeze(-1)).squeeze()
    return torch.argsort(-sc)


from newsrecommendationsystem_b200.recommend import Recommender
rec = Recommender(m, [f"N{i}" for i in range(n_rows - 1)], table)


def graph_us(fn, n=200):
    """device-side latency: the call captured in a CUDA graph, replayed back to back"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timeit(g.replay, n)


for C in (37, 300, 4096):
    cand = torch.from_numpy(rng.integers(0, n_rows, C).astype(np.int32)).to(dev)
    offs = torch.tensor([0, C], dtype=torch.int64, device=dev)
    t_lat = timeit(lambda: ops.recommend_user(table, hist, cand, *w))

    def batched():
        u = ops.user_encoder_indexed(table, hist.view(1, 50), *w, mode=_lib.MODE_FP32)
        s = ops.score_csr(table, cand, offs, u)
        return torch.argsort(-s)
    t_b = timeit(batched)
    t_e = timeit(lambda: eager_reference(cand))
    hl, cl = hist.cpu().numpy(), cand.cpu().numpy()
    t_rec = timeit(lambda: rec.recommend_rows(hl, cl))
    t_graph = graph_us(lambda: ops.recommend_user(table, hist, cand, *w))
    out[f"C={C}"] = {"recommend_user_us": round(t_lat, 2), "recommender_host_indices_us": round(t_rec, 2),
                     "recommend_user_graph_replay_us": round(t_graph, 2), "batched_fp32_kernels_us": round(t_b, 2),
                     "torch_eager_reference_ops_us": round(t_e, 2)}
    print(C, out[f"C={C}"], flush=True)
print(json.dumps(out))
