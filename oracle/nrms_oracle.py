"""CPU oracle for the NRMS hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm for the path
`news encoder -> user encoder -> dot-product click predictor` (forward, backward,
cross-entropy with label 0, Adam/AdamW, the evaluate orchestration and the ranking
metrics).  It is the CHECKER for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.  The product package `newsrecommendationsystem_b200` never imports it and
fails loudly if its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the LIVE reference
modules (`/root/reference/src/model/NRMS`, torch CPU fp32 + autograd + torch.optim.Adam,
and `evaluate.calculate_single_user_metric`) run in the build container by
`tests/golden/make_golden.py`; the resulting fixtures are committed under
`tests/golden/` and `tests/test_oracle_golden.py` checks every function here against them.
The restatements of model/Exp1 and of recommend.py's single-user arithmetic (end of this file) are pinned the same way
by `tests/golden/make_golden_exp1.py` (live reference `Exp1` / `NRMS` modules) and `tests/test_exp1_recommend.py`.

All `file:line` citations are relative to the reference tree (`/root/reference/`).
The arithmetic itself lives in third-party, un-vendored PyTorch (requirements.txt:1,
unpinned; container has 2.11.0) and scikit-learn (`roc_auc_score`, requirements.txt:7,
unpinned; container 1.9.0): their published semantics are restated below.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------------------
# parameter naming: the reference state_dict keys
# --------------------------------------------------------------------------------------
NEWS = "news_encoder"
USER = "user_encoder"


def enc_keys(prefix: str):
    m = f"{prefix}.multihead_self_attention"
    a = f"{prefix}.additive_attention"
    return dict(
        Wq=f"{m}.W_Q.weight", bq=f"{m}.W_Q.bias",
        Wk=f"{m}.W_K.weight", bk=f"{m}.W_K.bias",
        Wv=f"{m}.W_V.weight", bv=f"{m}.W_V.bias",
        Wa=f"{a}.linear.weight", ba=f"{a}.linear.bias",
        qa=f"{a}.attention_query_vector",
    )


EMB_KEY = "news_encoder.word_embedding.weight"


def ln_keys(prefix: str):
    """config-5 variant (builder-defined, no reference code: README.md:105-112): nn.LayerNorm(300) on the
    self-attention context of an encoder.  Present in `params` only for the variant."""
    return dict(ln_g=f"{prefix}.layer_norm.weight", ln_b=f"{prefix}.layer_norm.bias")


def enc_params(params: dict, prefix: str) -> dict:
    p = {k: params[v] for k, v in enc_keys(prefix).items()}
    for k, v in ln_keys(prefix).items():
        if v in params:
            p[k] = params[v]
    return p


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def embedding_gather(E: np.ndarray, tokens: np.ndarray) -> np.ndarray:
    """`self.word_embedding(news["title"])` -- src/model/NRMS/news_encoder.py:38.
    Pure row gather; bit-exact copy of the fp32 rows (padding_idx only affects grads)."""
    return E[tokens]


def mhsa_forward(x, p, num_heads, length=None):
    """MultiHeadSelfAttention.forward with K=V=Q; `length` (int [B]) is the optional mask branch (:60-68:
    attn_mask[b,h,i,j] = j < length[b]; :18-19: exp(scores) * attn_mask).  With length=None:
    (src/model/general/attention/multihead_self.py:46-76) followed by
    ScaledDotProductAttention.forward (:15-23).

    x: [B,S,D].  Q/K/V = x W^T + b (nn.Linear, :53-58); heads are contiguous 20-wide
    column slices (view(B,-1,H,d).transpose(1,2)); scores = QK^T / sqrt(d) (:16);
    exp WITHOUT max subtraction (:17); attn = e / (sum_j e + 1e-8) (:20); ctx = attn V
    (:22); heads merged back to [B,S,D] with NO output projection (:74-75)."""
    B, S, D = x.shape
    d = D // num_heads
    dt = x.dtype
    q = x @ p["Wq"].T + p["bq"]
    k = x @ p["Wk"].T + p["bk"]
    v = x @ p["Wv"].T + p["bv"]
    qh = q.reshape(B, S, num_heads, d).transpose(0, 2, 1, 3)
    kh = k.reshape(B, S, num_heads, d).transpose(0, 2, 1, 3)
    vh = v.reshape(B, S, num_heads, d).transpose(0, 2, 1, 3)
    s = (qh @ kh.transpose(0, 1, 3, 2)) / dt.type(np.sqrt(d))
    e = np.exp(s)
    if length is not None:
        mask = (np.arange(S)[None, :] < np.asarray(length).reshape(-1, 1)).astype(dt)      # [B,S] over key positions
        e = e * mask[:, None, None, :]
    attn = e / (e.sum(-1, keepdims=True) + dt.type(1e-8))
    ctx = attn @ vh
    out = ctx.transpose(0, 2, 1, 3).reshape(B, S, D)
    cache = dict(x=x, qh=qh, kh=kh, vh=vh, attn=attn)
    return out, cache


def additive_forward(c, p):
    """AdditiveAttention.forward (src/model/general/attention/additive.py:27-53):
    temp = tanh(linear(c)) (:35); weights = softmax(temp @ q, dim=1) (:37-39, torch's
    stable softmax); target = bmm(weights[:,None,:], c) (:51-52)."""
    t = np.tanh(c @ p["Wa"].T + p["ba"])
    s = t @ p["qa"]
    s = s - s.max(axis=1, keepdims=True)
    e = np.exp(s)
    w = e / e.sum(axis=1, keepdims=True)
    out = np.einsum("bs,bsd->bd", w, c)
    return out, dict(c=c, t=t, w=w)


def encoder_forward(x, p, num_heads, mask2=None):
    """MHSA -> (dropout #2, news encoder only) -> additive pooling.
    news: src/model/NRMS/news_encoder.py:41-47; user: src/model/NRMS/user_encoder.py:23-26.
    `mask2` is an explicit multiplicative dropout mask (already scaled by 1/(1-p))."""
    c, cm = mhsa_forward(x, p, num_heads)
    if mask2 is not None:
        c = c * mask2
    cl = None
    if "ln_g" in p:                      # config-5 variant: c = LayerNorm(dropout(MHSA(x)))
        c, cl = layernorm_forward(c, p["ln_g"], p["ln_b"])
    out, ca = additive_forward(c, p)
    return out, dict(mhsa=cm, add=ca, mask2=mask2, ln=cl)


def news_encoder_forward(params, tokens, num_heads=15, mask1=None, mask2=None):
    """NewsEncoder.forward (src/model/NRMS/news_encoder.py:27-48).  tokens: int [B,L].
    Dropout (F.dropout, :38-45) is expressed with explicit masks (None = eval mode)."""
    x = embedding_gather(params[EMB_KEY], tokens)
    if mask1 is not None:
        x = x * mask1
    out, cache = encoder_forward(x, enc_params(params, NEWS), num_heads, mask2)
    cache["tokens"] = tokens
    cache["mask1"] = mask1
    return out, cache


def user_encoder_forward(params, clicked_news_vector, num_heads=15):
    """UserEncoder.forward (src/model/NRMS/user_encoder.py:15-26); no dropout."""
    return encoder_forward(clicked_news_vector, enc_params(params, USER), num_heads)


def click_score(candidate_news_vector, user_vector):
    """DotProductClickPredictor.forward (src/model/general/click_predictor/dot_product.py:17-18):
    bmm([B,C,X],[B,X,1]).squeeze(-1) -> [B,C]."""
    return np.einsum("bcx,bx->bc", candidate_news_vector, user_vector)


def nrms_forward(params, cand_tokens, clicked_tokens, num_heads=15, masks=None):
    """NRMS.forward (src/model/NRMS/__init__.py:19-48).
    cand_tokens: int [B,1+K,L]; clicked_tokens: int [B,N,L] (the reference passes lists of
    per-position dicts and stacks on dim 1, :38-42; the news encoder is position-wise so
    one batched call over B*(1+K+N) titles is the same function)."""
    B, C, L = cand_tokens.shape
    N = clicked_tokens.shape[1]
    toks = np.concatenate([cand_tokens, clicked_tokens], axis=1).reshape(B * (C + N), L)
    m1 = m2 = None
    if masks is not None:
        m1, m2 = masks
    nv, cn = news_encoder_forward(params, toks, num_heads, m1, m2)
    nv = nv.reshape(B, C + N, -1)
    cand_v, clicked_v = nv[:, :C], nv[:, C:]
    uv, cu = user_encoder_forward(params, clicked_v, num_heads)
    logits = click_score(cand_v, uv)
    return logits, dict(news=cn, user=cu, cand_v=cand_v, clicked_v=clicked_v, uv=uv,
                        shapes=(B, C, N, L))


def cross_entropy_label0(logits):
    """nn.CrossEntropyLoss()(y_pred, zeros) (src/train.py:126,205-206): mean over the
    batch of -log_softmax(logits)[:, 0].  Returns (loss, dlogits)."""
    m = logits.max(axis=1, keepdims=True)
    z = logits - m
    lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
    logp = z - lse
    B = logits.shape[0]
    loss = -logp[:, 0].mean()
    g = np.exp(logp)
    g[:, 0] -= 1
    return logits.dtype.type(loss), (g / logits.dtype.type(B)).astype(logits.dtype)


# --------------------------------------------------------------------------------------
# backward (what torch autograd computes for the modules above)
# --------------------------------------------------------------------------------------
def additive_backward(dout, cache, p):
    c, t, w = cache["c"], cache["t"], cache["w"]
    dc = w[:, :, None] * dout[:, None, :]
    dw = np.einsum("bd,bsd->bs", dout, c)
    ds = w * (dw - (w * dw).sum(axis=1, keepdims=True))
    dqa = np.einsum("bs,bsq->q", ds, t)
    dt = ds[:, :, None] * p["qa"][None, None, :]
    du = dt * (1 - t * t)
    dWa = np.einsum("bsq,bsd->qd", du, c)
    dba = du.sum(axis=(0, 1))
    dc = dc + du @ p["Wa"]
    return dc, dict(Wa=dWa, ba=dba, qa=dqa)


def mhsa_backward(dout, cache, p, num_heads):
    x, qh, kh, vh, attn = (cache[k] for k in ("x", "qh", "kh", "vh", "attn"))
    B, S, D = x.shape
    d = D // num_heads
    dt = x.dtype
    dctx = dout.reshape(B, S, num_heads, d).transpose(0, 2, 1, 3)
    dvh = attn.transpose(0, 1, 3, 2) @ dctx
    dattn = dctx @ vh.transpose(0, 1, 3, 2)
    # attn = e/(sum e + eps)  =>  ds = attn * (dattn - sum_j attn*dattn)
    ds = attn * (dattn - (attn * dattn).sum(-1, keepdims=True))
    ds = ds / dt.type(np.sqrt(d))
    dqh = ds @ kh
    dkh = ds.transpose(0, 1, 3, 2) @ qh
    merge = lambda a: a.transpose(0, 2, 1, 3).reshape(B, S, D)
    dq, dk, dv = merge(dqh), merge(dkh), merge(dvh)
    x2 = x.reshape(B * S, D)
    g = dict(
        Wq=dq.reshape(-1, D).T @ x2, bq=dq.sum(axis=(0, 1)),
        Wk=dk.reshape(-1, D).T @ x2, bk=dk.sum(axis=(0, 1)),
        Wv=dv.reshape(-1, D).T @ x2, bv=dv.sum(axis=(0, 1)),
    )
    dx = dq @ p["Wq"] + dk @ p["Wk"] + dv @ p["Wv"]
    return dx, g


def encoder_backward(dout, cache, p, num_heads):
    dc, ga = additive_backward(dout, cache["add"], p)
    if cache.get("ln") is not None:
        dc, dg, db = layernorm_backward(dc, cache["ln"])
        ga.update(ln_g=dg, ln_b=db)
    if cache["mask2"] is not None:
        dc = dc * cache["mask2"]
    dx, gm = mhsa_backward(dc, cache["mhsa"], p, num_heads)
    gm.update(ga)
    return dx, gm


def nrms_backward(dlogits, cache, params, num_heads=15):
    """Gradients of every reference parameter (state_dict keys) for NRMS.forward.
    Embedding grad is dense with row 0 (padding_idx) zeroed -- news_encoder.py:15-20."""
    B, C, N, L = cache["shapes"]
    cand_v, uv = cache["cand_v"], cache["uv"]
    dcand = dlogits[:, :, None] * uv[:, None, :]
    duv = np.einsum("bc,bcx->bx", dlogits, cand_v)
    pu = enc_params(params, USER)
    dclicked, gu = encoder_backward(duv, cache["user"], pu, num_heads)
    dnv = np.concatenate([dcand, dclicked], axis=1).reshape(B * (C + N), -1)
    pn = enc_params(params, NEWS)
    dx, gn = encoder_backward(dnv, cache["news"], pn, num_heads)
    if cache["news"]["mask1"] is not None:
        dx = dx * cache["news"]["mask1"]
    E = params[EMB_KEY]
    dE = np.zeros_like(E)
    np.add.at(dE, cache["news"]["tokens"].reshape(-1), dx.reshape(-1, E.shape[1]))
    dE[0] = 0  # padding_idx=0
    grads = {EMB_KEY: dE}
    for k, name in enc_keys(NEWS).items():
        grads[name] = gn[k]
    for k, name in enc_keys(USER).items():
        grads[name] = gu[k]
    for prefix, g in ((NEWS, gn), (USER, gu)):
        for k, name in ln_keys(prefix).items():
            if k in g:
                grads[name] = g[k]
    return grads


# --------------------------------------------------------------------------------------
# optimizer (torch.optim.Adam / AdamW single-tensor semantics, torch 2.x)
# --------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8,
              weight_decay=0.0, decoupled=False):
    """torch.optim.Adam(lr=1e-4) as constructed at src/train.py:127-128 (defaults
    betas=(0.9,0.999), eps=1e-8, weight_decay=0, dense grads).  `decoupled=True` is
    torch.optim.AdamW (p *= 1 - lr*wd before the update).  `step` is 1-based.
    denom = sqrt(v)/sqrt(1-beta2^t) + eps ; p -= lr/(1-beta1^t) * m/denom."""
    f = p.dtype.type
    if weight_decay != 0.0:
        if decoupled:
            p = p * f(1 - lr * weight_decay)
        else:
            g = g + f(weight_decay) * p
    m = f(beta1) * m + f(1 - beta1) * g
    v = f(beta2) * v + f(1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p, m, v


def cosine_lr(base_lr, step, total_steps, eta_min=0.0):
    """torch.optim.lr_scheduler.CosineAnnealingLR closed form (config-5 variant; the
    reference has no scheduler code -- README.md:112 only; builder-defined)."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * step / total_steps)) / 2


def train_step(params, opt_state, cand_tokens, clicked_tokens, step, num_heads=15,
               lr=1e-4, weight_decay=0.0, decoupled=False):
    """One iteration of the hot loop, src/train.py:202-206,227-233 (eval-mode dropout)."""
    logits, cache = nrms_forward(params, cand_tokens, clicked_tokens, num_heads)
    loss, dlogits = cross_entropy_label0(logits)
    grads = nrms_backward(dlogits, cache, params, num_heads)
    new_p, new_s = {}, {}
    for k in params:
        m, v = opt_state.get(k, (np.zeros_like(params[k]), np.zeros_like(params[k])))
        p2, m2, v2 = adam_step(params[k], grads[k], m, v, step, lr=lr,
                               weight_decay=weight_decay, decoupled=decoupled)
        new_p[k], new_s[k] = p2, (m2, v2)
    return loss, logits, grads, new_p, new_s


# --------------------------------------------------------------------------------------
# LayerNorm variant (config 5; builder-defined): nn.LayerNorm(300) on the
# MHSA output before additive attention.
# --------------------------------------------------------------------------------------
def layernorm_forward(x, gamma, beta, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    rstd = 1 / np.sqrt(var + x.dtype.type(eps))
    xh = (x - mu) * rstd
    return xh * gamma + beta, dict(xh=xh, rstd=rstd, gamma=gamma)


def layernorm_backward(dy, cache):
    xh, rstd, gamma = cache["xh"], cache["rstd"], cache["gamma"]
    dgamma = (dy * xh).reshape(-1, xh.shape[-1]).sum(0)
    dbeta = dy.reshape(-1, xh.shape[-1]).sum(0)
    dxh = dy * gamma
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------
# ranking metrics (src/evaluate.py:24-42,160-168) -- fp64 like numpy/sklearn
# --------------------------------------------------------------------------------------
def _order_desc(y_score):
    """`np.argsort(y_score)[::-1]` (src/evaluate.py:25,39).  numpy's default sort is not
    stable, so the reference leaves the relative order of EXACT ties unspecified; the
    oracle (and the CUDA kernel) fix it as reversed-stable: descending score, and among
    equal scores descending index."""
    return np.argsort(np.asarray(y_score), kind="stable")[::-1]


def dcg_score(y_true, y_score, k=10):
    order = _order_desc(y_score)
    yt = np.take(np.asarray(y_true), order[:k])
    gains = 2.0 ** yt - 1
    discounts = np.log2(np.arange(len(yt)) + 2)
    return float(np.sum(gains / discounts))


def ndcg_score(y_true, y_score, k=10):
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true, k)


def mrr_score(y_true, y_score):
    order = _order_desc(y_score)
    yt = np.take(np.asarray(y_true), order)
    rr = yt / (np.arange(len(yt)) + 1)
    return float(np.sum(rr) / np.sum(yt))


def auc_score(y_true, y_score):
    """sklearn.metrics.roc_auc_score for binary labels (src/evaluate.py:2,162): area under
    the ROC curve by the trapezoidal rule == Mann-Whitney U / (P*N) with ties counted 1/2.
    Raises ValueError when only one class is present (the reference maps it to NaN, :167)."""
    yt = np.asarray(y_true)
    ys = np.asarray(y_score, dtype=np.float64)
    pos = ys[yt == 1]
    neg = ys[yt == 0]
    if len(pos) == 0 or len(neg) == 0:
        raise ValueError("Only one class present in y_true.")
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return float((gt + 0.5 * eq) / (len(pos) * len(neg)))


def single_user_metric(y_true, y_score):
    """calculate_single_user_metric (src/evaluate.py:160-168)."""
    try:
        return [auc_score(y_true, y_score), mrr_score(y_true, y_score),
                ndcg_score(y_true, y_score, 5), ndcg_score(y_true, y_score, 10)]
    except ValueError:
        return [np.nan] * 4


# --------------------------------------------------------------------------------------
# evaluate orchestration (src/evaluate.py:185-272) on integer tables
# --------------------------------------------------------------------------------------
def first_occurrence_rows(news_ids):
    """news2vector is filled `if id not in news2vector` (src/evaluate.py:197-201): the FIRST
    row carrying an id wins.  Returns, for each row, the row index that owns its id."""
    first = {}
    owner = np.empty(len(news_ids), dtype=np.int64)
    for i, nid in enumerate(news_ids.tolist()):
        owner[i] = first.setdefault(nid, i)
    return owner


def build_history(clicked_rows_list, num_clicked=50, pad_row=-1):
    """UserDataset.__getitem__ (src/evaluate.py:111-124): keep the FIRST `num_clicked`
    clicks, LEFT-pad with PADDED_NEWS.  `clicked_rows_list` is a list of int sequences."""
    out = np.full((len(clicked_rows_list), num_clicked), pad_row, dtype=np.int64)
    for i, h in enumerate(clicked_rows_list):
        h = list(h)[:num_clicked]
        if h:
            out[i, num_clicked - len(h):] = h
    return out


def evaluate_pipeline(params, news_tokens, hist_rows, cand_offsets, cand_rows, labels,
                      num_heads=15, news_owner=None, max_count=None, batch=2048):
    """Restatement of evaluate() (src/evaluate.py:171-272) on tensors.
    news_tokens: int [Nn,L]; hist_rows: int [I,N] rows into the news table, -1 = PADDED_NEWS
    (a literal zero vector, :203-204); cand_offsets: int [I+1] CSR; cand_rows/labels: int [sumC].
    `max_count` reproduces the off-by-one at :247-249 (processes max_count-1 impressions).
    Returns (means[4], per_impression[I',4], scores[sumC'], news_table, user_vectors)."""
    Nn = news_tokens.shape[0]
    D = params[EMB_KEY].shape[1]
    table = np.zeros((Nn + 1, D), dtype=params[EMB_KEY].dtype)  # last row = PADDED_NEWS
    for s in range(0, Nn, batch):
        e = min(s + batch, Nn)
        table[s:e] = news_encoder_forward(params, news_tokens[s:e], num_heads)[0]
    if news_owner is not None:
        table[:Nn] = table[news_owner]
    I = hist_rows.shape[0]
    uvec = np.zeros((I, D), dtype=table.dtype)
    hr = np.where(hist_rows < 0, Nn, hist_rows)
    for s in range(0, I, batch):
        uvec[s:s + batch] = user_encoder_forward(params, table[hr[s:s + batch]], num_heads)[0]
    n_imp = I if max_count is None else min(I, max_count - 1)
    scores = np.zeros(int(cand_offsets[n_imp]), dtype=table.dtype)
    per = np.zeros((n_imp, 4), dtype=np.float64)
    for i in range(n_imp):
        a, b = int(cand_offsets[i]), int(cand_offsets[i + 1])
        sc = table[cand_rows[a:b]] @ uvec[i]
        scores[a:b] = sc
        per[i] = single_user_metric(labels[a:b], sc.tolist())
    with np.errstate(all="ignore"):
        means = np.nanmean(per, axis=0)
    return means, per, scores, table, uvec


# --------------------------------------------------------------------------------------
# model/Exp1 and the single-user path of recommend.py
# (pinned by tests/golden/make_golden_exp1.py against the live reference modules)
# --------------------------------------------------------------------------------------
def _block_params(params, m_prefix, a_prefix):
    return dict(
        Wq=params[f"{m_prefix}.W_Q.weight"], bq=params[f"{m_prefix}.W_Q.bias"],
        Wk=params[f"{m_prefix}.W_K.weight"], bk=params[f"{m_prefix}.W_K.bias"],
        Wv=params[f"{m_prefix}.W_V.weight"], bv=params[f"{m_prefix}.W_V.bias"],
        Wa=params[f"{a_prefix}.linear.weight"], ba=params[f"{a_prefix}.linear.bias"],
        qa=params[f"{a_prefix}.attention_query_vector"])


def element_encoder_forward(E, W, b, element):
    """ElementEncoder.forward (src/model/Exp1/news_encoder.py:43-44): F.relu(self.linear(self.embedding(element)))."""
    return np.maximum(E[element] @ W.T + b, 0).astype(E.dtype)


def exp1_text_encoder_forward(params, name, text, num_heads=15):
    """TextEncoder.forward in eval mode (src/model/Exp1/news_encoder.py:20-34): the NRMS news-encoder block over the
    shared word embedding with the text encoder's own attention weights."""
    pre = f"news_encoder.text_encoders.{name}"
    x = embedding_gather(params[f"{pre}.word_embedding.weight"], text)
    out, _ = encoder_forward(x, _block_params(params, f"{pre}.multihead_self_attention", f"{pre}.additive_attention"),
                             num_heads)
    return out


def exp1_news_encoder_forward(params, news, attributes, num_heads=15):
    """Exp1 NewsEncoder.forward (src/model/Exp1/news_encoder.py:83-111): text vectors + element vectors, stacked on
    dim 1 and pooled by the final additive attention (a single vector is returned as is, :106-107)."""
    texts = [a for a in ("title", "abstract") if a in attributes]
    elems = [a for a in ("category", "subcategory") if a in attributes]
    vecs = [exp1_text_encoder_forward(params, a, news[a], num_heads) for a in texts]
    for a in elems:
        pre = f"news_encoder.element_encoders.{a}"
        vecs.append(element_encoder_forward(params[f"{pre}.embedding.weight"], params[f"{pre}.linear.weight"],
                                            params[f"{pre}.linear.bias"], news[a]))
    if len(vecs) == 1:
        return vecs[0]
    pf = "news_encoder.final_attention"
    out, _ = additive_forward(np.stack(vecs, axis=1),
                              dict(Wa=params[f"{pf}.linear.weight"], ba=params[f"{pf}.linear.bias"],
                                   qa=params[f"{pf}.attention_query_vector"]))
    return out


def exp1_user_encoder_forward(params, clicked_news_vector, num_heads=15):
    """Exp1 UserEncoder.forward (src/model/Exp1/user_encoder.py:15-31): the NRMS user-encoder block over
    user_vector + position_embedding."""
    x = clicked_news_vector + params["user_encoder.position_embedding"][None]
    out, _ = encoder_forward(x, _block_params(params, "user_encoder.multihead_self_attention",
                                              "user_encoder.additive_attention"), num_heads)
    return out


def exp1_forward(params, news, n_cand, attributes, num_heads=15):
    """Exp1.forward (src/model/Exp1/__init__.py:15-46).  news: {attribute: [B, 1+K+N, ...]} (candidates first)."""
    B, T = news["title"].shape[:2]
    flat = {a: news[a].reshape(B * T, *news[a].shape[2:]) for a in attributes}
    vec = exp1_news_encoder_forward(params, flat, attributes, num_heads).reshape(B, T, -1)
    uv = exp1_user_encoder_forward(params, vec[:, n_cand:], num_heads)
    return click_score(vec[:, :n_cand], uv)


def recommend_user(params, table, hist_rows, cand_rows, num_heads=15):
    """The single target user of recommend.py: user vector from the stacked cache rows (src/recommend.py:264-279),
    get_prediction over the impression's candidates (:301-315), y = (score + 1) / 2 in python floats (:338) and
    np.argsort(-y) (:339).  numpy's default sort leaves the order of EXACTLY equal scores unspecified; the order is
    defined here as stable (ascending candidate position), like the metric order in `_order_desc`."""
    uv, _ = user_encoder_forward(params, table[np.asarray(hist_rows)][None].astype(np.float32), num_heads)
    scores = table[np.asarray(cand_rows, dtype=np.int64)] @ uv[0] if len(cand_rows) else np.zeros(0, np.float32)
    y = (scores.astype(np.float64) + 1.0) / 2.0
    return uv[0], y, np.argsort(-y, kind="stable")
