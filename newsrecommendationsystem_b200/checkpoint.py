"""Checkpoint and news-vector cache interchange with the reference.

* `src/train.py:266-277` writes `{"model_state_dict", "optimizer_state_dict", "step", "early_stop_value"}` with
  torch.save; `src/evaluate.py:280-289` / `src/recommend.py:368-381` read `model_state_dict` back.  The B200 module
  keeps the reference's parameter names and shapes, and `FusedAdam.state_dict()` is laid out like
  torch.optim.Adam's, so the same file serves both implementations in both directions.
* `src/recommend.py:211-243` caches `news2vector.pt`: a dict news id -> vector (+ `PADDED_NEWS` zeros).  The
  device-resident evaluate path keeps a table [N_news + 1, 300] whose last row is PADDED_NEWS; the two helpers
  below convert between the forms (first occurrence of an id wins, as in `src/evaluate.py:197-201`).
"""
from __future__ import annotations

import torch


def save_checkpoint(path, model, optimizer, step, early_stop_value):
    """The reference's checkpoint dict (src/train.py:266-277); tensors are moved to the CPU first."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    osd = optimizer.state_dict()
    for st in osd["state"].values():
        for k, v in st.items():
            if torch.is_tensor(v):
                st[k] = v.detach().cpu()
    ckpt = {"model_state_dict": sd, "optimizer_state_dict": osd, "step": step, "early_stop_value": early_stop_value}
    ne = getattr(model, "news_encoder", None)
    if ne is not None and hasattr(ne, "_dropout_calls"):
        # position of the in-kernel Philox dropout stream: a resumed run continues it instead of replaying it from 0
        # (an extra key the reference's loader ignores: it reads the four keys above only, train.py:144-153)
        ckpt["b200_dropout_state"] = {"seed": int(ne.dropout_seed), "offset": int(ne._dropout_calls)}
    torch.save(ckpt, path)


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """Load a checkpoint written by the reference (or by `save_checkpoint`) -> (step, early_stop_value).
    Parameters are copied IN PLACE (views into FusedAdam's flat buffer stay valid)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"])
    if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    ds = ckpt.get("b200_dropout_state")
    ne = getattr(model, "news_encoder", None)
    if ds is not None and ne is not None and hasattr(ne, "_dropout_calls"):
        ne.dropout_seed, ne._dropout_calls = int(ds["seed"]), int(ds["offset"])
    return ckpt.get("step"), ckpt.get("early_stop_value")


def news2vector_from_table(news_ids, table):
    """ids (list of str, row order of the token table) + table [N + 1, X] -> the reference's news2vector dict."""
    out = {}
    for i, nid in enumerate(news_ids):
        if nid not in out:                       # first occurrence wins (src/evaluate.py:197-201)
            out[nid] = table[i]
    out["PADDED_NEWS"] = torch.zeros_like(table[0])
    return out


def table_from_news2vector(news2vector, news_ids=None, device=None):
    """Reference news2vector dict -> (ids, table [N + 1, X] with the PADDED_NEWS zero row last)."""
    ids = list(news_ids) if news_ids is not None else [k for k in news2vector if k != "PADDED_NEWS"]
    rows = [news2vector[i] for i in ids]
    first = rows[0]
    table = torch.stack([r.to(first.device) for r in rows] + [torch.zeros_like(first)])
    return ids, (table.to(device) if device is not None else table)
