"""GPU parity tests of the table path with the additive pooling fused into the attention kernel (K1f,
csrc/k1f_attn_pool.cu, option "fused_pool") and of the overflow-free softmax form (large attention scores) in every
tensor-mode attention kernel: K1g (table path, default), K1f, K1 v6 (per-sequence projection).

Reference math: src/model/general/attention/multihead_self.py:15-23 (exp / (sum + 1e-8), no max subtraction) and
src/model/general/attention/additive.py:27-53, composed by src/model/NRMS/user_encoder.py:15-26 and
src/model/NRMS/news_encoder.py:27-48.  Tolerance: <= 1e-3 max row-wise relative L2 in tensor mode (north_star).
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2_rows
from oracle import nrms_oracle as O
from test_gpu_parity import Cfg, make_model, t, TOL_VEC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    from newsrecommendationsystem_b200 import _lib
    return _lib.load()


def _user_case(n_users, n_rows, seed=None):
    rng = np.random.default_rng(n_users if seed is None else seed)
    table = (rng.standard_normal((n_rows + 1, 300)) * 0.3).astype(np.float32)
    table[n_rows] = 0                                     # PADDED_NEWS
    rows = rng.integers(0, n_rows, size=(n_users, 50))
    rows[0, :50 - min(49, n_users)] = n_rows              # left-padded history
    if n_users > 9:
        rows[9] = n_rows                                  # empty history
        rows[3] = rows[3, 0]                              # one news repeated 50 times
    assert n_users * 50 >= 8 * (n_rows + 1)               # the size rule that selects the table path
    return table, rows


@pytest.mark.parametrize("safe", [-1, 0, 1])
@pytest.mark.parametrize("n_users,n_rows", [(1, 3), (13, 60), (67, 300), (149, 900), (200, 1000), (2500, 4001)])
def test_user_encoder_fused_pool(dev, lib, golden_sd, n_users, n_rows, safe):
    """K1f (one launch: attention over gathered q|k|v rows + additive pooling) vs the oracle, vs the two-kernel table
    path (K1g -> HBM -> K2) and vs the per-user projection (K1 v6 + K2).  safe: -1 = the kernel decides from the score
    bound, 0 = plain 2^s, 1 = row-shifted form (identical math, must agree to rounding)."""
    table, rows = _user_case(n_users, n_rows)
    ref, _ = O.user_encoder_forward(golden_sd, table[rows])
    m = make_model(golden_sd, dev, "tf32")
    tb, ix = t(table, dev), t(rows.astype(np.int32), dev)
    with torch.no_grad():
        try:
            assert lib.nrms_set_option(b"attn_safe_softmax", safe) == 0
            assert lib.nrms_set_option(b"fused_pool", 1) == 0
            a = m.user_encoder.forward_indexed(tb, ix)
            a2 = m.user_encoder.forward_indexed(tb, ix)
            assert lib.nrms_set_option(b"fused_pool", 0) == 0
            b = m.user_encoder.forward_indexed(tb, ix)
            lib.nrms_set_option(b"user_table_attn", 0)
            c = m.user_encoder.forward_indexed(tb, ix)
        finally:
            lib.nrms_set_option(b"user_table_attn", 1)
            lib.nrms_set_option(b"fused_pool", 0)
            lib.nrms_set_option(b"attn_safe_softmax", -1)
    assert torch.isfinite(a).all()
    assert torch.equal(a, a2)                             # fixed summation order: bit-identical reruns
    ea, eb, ec = (rel_l2_rows(x.cpu().numpy(), ref) for x in (a, b, c))
    print(f"users {n_users}x{n_rows} safe={safe}: K1f {ea:.2e}  K1g+K2 {eb:.2e}  K1v6+K2 {ec:.2e}")
    assert ea < TOL_VEC["tf32"] and eb < TOL_VEC["tf32"] and ec < TOL_VEC["tf32"]
    assert not torch.equal(a, c)                          # different kernels really ran
    assert rel_l2_rows(a.cpu().numpy(), b.cpu().numpy().astype(np.float64)) < TOL_VEC["tf32"]


@pytest.mark.parametrize("safe", [-1, 1])
@pytest.mark.parametrize("n_titles,num_words", [(2, 3), (37, 80), (445, 1001), (777, 401), (6000, 2001)])
def test_news_encoder_fused_pool(dev, lib, golden_sd, n_titles, num_words, safe):
    """The same kernel at S = 20 (three titles per context tile), token ids int64 into the projected embedding table."""
    from newsrecommendationsystem_b200 import synthetic
    sd = dict(golden_sd)
    rng = np.random.default_rng(num_words)
    emb = rng.standard_normal((num_words, 300)).astype(np.float32)
    emb[0] = 0
    sd["news_encoder.word_embedding.weight"] = emb

    class C2(Cfg):
        pass
    C2.num_words = num_words
    toks = synthetic.make_news(n_titles, num_words=num_words, seed=n_titles)
    if n_titles > 11:
        toks[11] = 0                                        # an all-padding title
    assert n_titles * 20 >= 8 * num_words
    ref, _ = O.news_encoder_forward(sd, toks)
    m = make_model(sd, dev, "tf32", cfg=C2)
    with torch.no_grad():
        try:
            assert lib.nrms_set_option(b"attn_safe_softmax", safe) == 0
            lib.nrms_set_option(b"fused_pool", 1)
            a = m.get_news_vector({"title": torch.from_numpy(toks)})
            lib.nrms_set_option(b"fused_pool", 0)
            b = m.get_news_vector({"title": torch.from_numpy(toks)})
        finally:
            lib.nrms_set_option(b"fused_pool", 0)
            lib.nrms_set_option(b"attn_safe_softmax", -1)
    ea, eb = rel_l2_rows(a.cpu().numpy(), ref), rel_l2_rows(b.cpu().numpy(), ref)
    print(f"news {n_titles}x{num_words} safe={safe}: K1f {ea:.2e}  K1g+K2 {eb:.2e}")
    assert torch.isfinite(a).all()
    assert ea < TOL_VEC["tf32"] and eb < TOL_VEC["tf32"]
    assert rel_l2_rows(a.cpu().numpy(), b.cpu().numpy().astype(np.float64)) < TOL_VEC["tf32"]


def _scale_qk(sd, prefix, x, target_nats):
    """Scale W_Q, b_Q, W_K, b_K of one encoder so that the largest attention logit q.k/sqrt(20) over the inputs x
    [n, S, 300] becomes `target_nats` (a trained / GloVe-initialised checkpoint can have such scores; random init has
    <= ~7).  Returns (new state dict, achieved maximum)."""
    p = O.enc_params(sd, prefix)
    q = x @ p["Wq"].T + p["bq"]
    k = x @ p["Wk"].T + p["bk"]
    n, S, _ = x.shape
    qh = q.reshape(n, S, 15, 20).transpose(0, 2, 1, 3)
    kh = k.reshape(n, S, 15, 20).transpose(0, 2, 1, 3)
    smax = float((qh @ kh.transpose(0, 1, 3, 2)).max() / np.sqrt(20.0))
    f = np.float32(np.sqrt(target_nats / smax))
    out = dict(sd)
    for w in ("W_Q", "W_K"):
        for leaf in ("weight", "bias"):
            key = f"{prefix}.multihead_self_attention.{w}.{leaf}"
            out[key] = (sd[key] * f).astype(np.float32)
    return out, smax * float(f) ** 2


@pytest.mark.parametrize("target", [20.0, 40.0, 80.0])
@pytest.mark.parametrize("path", ["k1g", "k1f", "k1v6"])
def test_user_encoder_large_scores(dev, lib, golden_sd, target, path):
    """Attention logits of 20 / 40 / 80 nats: 2^s packed to fp16 would be inf from 11.09 on, the reference's fp32 exp
    (multihead_self.py:17) is finite to 88.  Every tensor-mode kernel must stay finite and follow the oracle.
    The relative tolerance grows with the score: the fp16 / tf32 operands carry 2^-11 relative error, i.e. an
    ABSOLUTE error of ~s * 5e-4 nats on a logit, so near-tied keys change weight by that fraction."""
    table, rows = _user_case(200, 1000, seed=77)
    sd, smax = _scale_qk(golden_sd, O.USER, table[rows].astype(np.float32), target)
    ref, _ = O.user_encoder_forward({k: np.asarray(v, dtype=np.float64) for k, v in sd.items()},
                                    table[rows].astype(np.float64))
    m = make_model(sd, dev, "tf32")
    tb, ix = t(table, dev), t(rows.astype(np.int32), dev)
    with torch.no_grad():
        try:
            if path == "k1f":
                lib.nrms_set_option(b"fused_pool", 1)
            if path == "k1v6":
                lib.nrms_set_option(b"user_table_attn", 0)
            a = m.user_encoder.forward_indexed(tb, ix)
            m32 = make_model(sd, dev, "fp32")
            f = m32.user_encoder.forward_indexed(tb, ix)
        finally:
            lib.nrms_set_option(b"user_table_attn", 1)
            lib.nrms_set_option(b"fused_pool", 0)
    assert torch.isfinite(a).all(), f"{path}: non-finite user vectors at {smax:.1f} nats"
    ea, ef = rel_l2_rows(a.cpu().numpy(), ref), rel_l2_rows(f.cpu().numpy(), ref)
    print(f"users large scores {smax:.1f} nats [{path}]: tensor {ea:.2e}  fp32 mode {ef:.2e}")
    assert ef < 1e-3
    assert ea < 1e-3 * max(1.0, 2.0 * smax)


@pytest.mark.parametrize("target", [20.0, 40.0])
@pytest.mark.parametrize("path", ["k1g", "k1f", "k1v6"])
def test_news_encoder_large_scores(dev, lib, golden_sd, target, path):
    from newsrecommendationsystem_b200 import synthetic
    num_words, n_titles = 401, 777
    sd = dict(golden_sd)
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((num_words, 300)).astype(np.float32)
    emb[0] = 0
    sd["news_encoder.word_embedding.weight"] = emb
    toks = synthetic.make_news(n_titles, num_words=num_words, seed=5)
    sd, smax = _scale_qk(sd, O.NEWS, emb[toks], target)
    ref, _ = O.news_encoder_forward({k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}, toks)
    m = make_model(sd, dev, "tf32")
    with torch.no_grad():
        try:
            if path == "k1f":
                lib.nrms_set_option(b"fused_pool", 1)
            if path == "k1v6":
                lib.nrms_set_option(b"news_table_attn", 0)
            a = m.get_news_vector({"title": torch.from_numpy(toks)})
        finally:
            lib.nrms_set_option(b"news_table_attn", 1)
            lib.nrms_set_option(b"fused_pool", 0)
    assert torch.isfinite(a).all(), f"{path}: non-finite news vectors at {smax:.1f} nats"
    ea = rel_l2_rows(a.cpu().numpy(), ref)
    print(f"news large scores {smax:.1f} nats [{path}]: tensor {ea:.2e}")
    assert ea < 1e-3 * max(1.0, 2.0 * smax)
