"""torch-CPU port of the reference's NRMS op sequence -- TEST/BENCH INFRASTRUCTURE ONLY.

Same header as oracle/nrms_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.  It exists because the fairest CPU
baseline for the reference is the reference's own arithmetic library (PyTorch CPU: oneDNN/MKL
addmm + bmm) driven with the reference's own op order; /root/reference is not present on the
GPU box, so the op sequence is restated here (file:line cited per function) and pinned against
the same golden fixtures as the numpy oracle (tests/test_oracle_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import nrms_oracle as O


def _p(params, prefix):
    return {k: torch.as_tensor(params[v]) for k, v in O.enc_keys(prefix).items()}


def mhsa(x, p, num_heads=15):
    """multihead_self.py:46-76 + :15-23 with the same view/transpose/contiguous sequence."""
    B = x.size(0)
    d = x.size(2) // num_heads
    q = F.linear(x, p["Wq"], p["bq"]).view(B, -1, num_heads, d).transpose(1, 2)
    k = F.linear(x, p["Wk"], p["bk"]).view(B, -1, num_heads, d).transpose(1, 2)
    v = F.linear(x, p["Wv"], p["bv"]).view(B, -1, num_heads, d).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)) / np.sqrt(d)
    scores = torch.exp(scores)
    attn = scores / (torch.sum(scores, dim=-1, keepdim=True) + 1e-8)
    ctx = torch.matmul(attn, v)
    return ctx.transpose(1, 2).contiguous().view(B, -1, num_heads * d)


def additive(c, p):
    """additive.py:27-53."""
    temp = torch.tanh(F.linear(c, p["Wa"], p["ba"]))
    w = F.softmax(torch.matmul(temp, p["qa"]), dim=1)
    return torch.bmm(w.unsqueeze(1), c).squeeze(1)


@torch.no_grad()
def news_vectors(params, tokens):
    """NRMS.get_news_vector, eval mode (NRMS/__init__.py:50-61, news_encoder.py:27-48)."""
    E = torch.as_tensor(params[O.EMB_KEY])
    x = F.embedding(torch.as_tensor(tokens), E, padding_idx=0)
    return additive(mhsa(x, _p(params, O.NEWS)), _p(params, O.NEWS))


@torch.no_grad()
def user_vectors(params, clicked):
    """NRMS.get_user_vector (NRMS/__init__.py:63-71, user_encoder.py:15-26)."""
    p = _p(params, O.USER)
    return additive(mhsa(torch.as_tensor(clicked), p), p)


@torch.no_grad()
def prediction(news_vector, user_vector):
    """NRMS.get_prediction (NRMS/__init__.py:73-84) incl. the reference's .tolist() (evaluate.py:260)."""
    return torch.bmm(news_vector.unsqueeze(0), user_vector.unsqueeze(0).unsqueeze(-1)).squeeze(-1).squeeze(0).tolist()


def train_step(params, cand_tokens, clicked_tokens, lr=1e-4, adam_state=None):
    """One iteration of the reference's training loop on the CPU (src/train.py:202-206,227-233): NRMS.forward
    (NRMS/__init__.py:19-48: every title through the news encoder, the 50 clicked vectors through the user encoder,
    dot-product scores), CrossEntropyLoss against label 0, loss.backward(), torch.optim.Adam(lr=1e-4).step().
    Eval-mode forward (dropout = identity; the reference's two F.dropout calls are elementwise and do not change the
    cost measurably).  cand_tokens int [B, 1+K, L], clicked_tokens int [B, N, L].  Returns (loss, parameters, optimizer)."""
    P = {k: torch.as_tensor(v).clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(P.values()), lr=lr) if adam_state is None else adam_state
    B, C1, L = cand_tokens.shape
    N = clicked_tokens.shape[1]
    with torch.enable_grad():
        toks = torch.as_tensor(np.concatenate([cand_tokens, clicked_tokens], axis=1)).reshape(B * (C1 + N), L)
        x = F.embedding(toks, P[O.EMB_KEY], padding_idx=0)
        pn = {k: P[v] for k, v in O.enc_keys(O.NEWS).items()}
        pu = {k: P[v] for k, v in O.enc_keys(O.USER).items()}
        vec = additive(mhsa(x, pn), pn).view(B, C1 + N, -1)
        user = additive(mhsa(vec[:, C1:], pu), pu)
        logits = torch.bmm(vec[:, :C1], user.unsqueeze(-1)).squeeze(-1)
        loss = F.cross_entropy(logits, torch.zeros(B, dtype=torch.long))
        opt.zero_grad()
        loss.backward()
        opt.step()
    return float(loss.item()), P, opt
