"""Pins the numpy oracle against fixtures produced by the live reference modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import nrms_oracle as O
from conftest import rel_l2_rows

ROWS = 48


def clip(a):
    return a[:ROWS] if a.ndim == 2 else a


def test_gather_bit_exact(golden, golden_sd):
    g = O.embedding_gather(golden_sd[O.EMB_KEY], golden["fwd/tokens"])
    assert g.dtype == np.float32
    assert np.array_equal(g.view(np.uint32), golden["fwd/gathered"].view(np.uint32))


def test_news_vectors(golden, golden_sd):
    nv, _ = O.news_encoder_forward(golden_sd, golden["fwd/tokens"])
    assert rel_l2_rows(nv, golden["fwd/news_vectors"]) < 2e-6


def test_user_vectors(golden, golden_sd):
    uv, _ = O.user_encoder_forward(golden_sd, golden["fwd/user_input"])
    assert rel_l2_rows(uv, golden["fwd/user_vectors"]) < 2e-6


def test_scores(golden):
    nv, uv = golden["fwd/news_vectors"], golden["fwd/user_vectors"]
    s = O.click_score(nv[None, :23], uv[None, 0])[0]
    np.testing.assert_allclose(s, golden["fwd/scores_single"], rtol=1e-5, atol=1e-6)
    sb = O.click_score(nv[:36].reshape(9, 4, 300), uv)
    np.testing.assert_allclose(sb, golden["fwd/scores_batched"], rtol=1e-5, atol=1e-6)


@pytest.fixture(scope="module")
def step1(golden, golden_sd):
    return O.train_step(golden_sd, {}, golden["train/cand"], golden["train/clicked"], step=1)


def test_train_forward_loss(golden, step1):
    loss, logits, *_ = step1
    np.testing.assert_allclose(logits, golden["train/logits"], rtol=2e-5, atol=2e-6)
    assert abs(float(loss) - float(golden["train/loss"])) < 2e-6


def test_train_grads(golden, step1):
    grads = step1[2]
    for k, g in grads.items():
        ref = golden["train/grad/" + k]
        got = g if "embedding" in k else clip(g)
        scale = np.abs(ref).max() + 1e-12
        # W_K.bias grads are analytically ~0 (row-softmax shift invariance): pure roundoff
        assert np.abs(got - ref).max() < 5e-5 * scale + 1e-8, k
    assert not grads[O.EMB_KEY][0].any()   # padding_idx row has zero grad


def close_adam(got, ref, k, atol=2e-7, lr=1e-4, gscale=1.0):
    """Adam's first steps are ill-conditioned where |g| ~ eps (update = lr*g/(|g|+eps)):
    roundoff-level grad differences move those elements by up to ~lr (the exact pin for the
    optimizer is test_adam_elementwise_exact).  Require >=80% of the elements inside `atol`,
    a mean error under 1e-6 and every element inside 2.5*lr."""
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert d.max() <= 2.5 * lr, k
    if gscale < 1e-5:
        return   # |g| within ~1000x of eps=1e-8: this tensor's Adam update amplifies roundoff
    assert (d <= atol).mean() >= 0.80, (k, float((d <= atol).mean()))
    assert d.mean() < 1e-6, (k, float(d.mean()))


def gs(golden, k):
    return float(np.abs(golden["train/grad/" + k]).mean())


def test_adam_elementwise_exact(golden, golden_sd):
    """adam_step on the reference's own grads reproduces torch.optim.Adam to 1 ulp-ish."""
    for k in golden_sd:
        if "embedding" in k:
            continue
        p, g = clip(golden_sd[k]), golden["train/grad/" + k]
        p1, _, _ = O.adam_step(p, g, np.zeros_like(p), np.zeros_like(p), 1)
        np.testing.assert_allclose(p1, golden["train/adam1/" + k], rtol=0, atol=1.5e-8, err_msg=k)
        pw, _, _ = O.adam_step(p, g, np.zeros_like(p), np.zeros_like(p), 1, weight_decay=0.01, decoupled=True)
        np.testing.assert_allclose(pw, golden["train/adamw1/" + k], rtol=0, atol=1.5e-8, err_msg=k)


def test_adam_two_steps(golden, golden_sd, step1):
    _, _, _, p1, s1 = step1
    for k in p1:
        close_adam(clip(p1[k]), golden["train/adam1/" + k], k, gscale=gs(golden, k))
    loss2, _, _, p2, _ = O.train_step(p1, s1, golden["train/cand"], golden["train/clicked"], step=2)
    assert abs(float(loss2) - float(golden["train/loss2"])) < 5e-6
    for k in p2:
        close_adam(clip(p2[k]), golden["train/adam2/" + k], k, atol=4e-7, gscale=gs(golden, k))


def test_adamw_step(golden, golden_sd):
    _, _, _, p1, _ = O.train_step(golden_sd, {}, golden["train/cand"], golden["train/clicked"], step=1,
                                  weight_decay=0.01, decoupled=True)
    for k in p1:
        close_adam(clip(p1[k]), golden["train/adamw1/" + k], k, gscale=gs(golden, k))


def test_metrics(golden):
    offs = golden["metric/offsets"]
    for i in range(len(offs) - 1):
        a, b = offs[i], offs[i + 1]
        got = O.single_user_metric(golden["metric/labels"][a:b].astype(np.int64),
                                   golden["metric/scores"][a:b].tolist())
        ref = golden["metric/results"][i]
        if np.isnan(ref).any():
            assert np.isnan(got).all()
            continue
        y = golden["metric/labels"][a:b]
        sc = golden["metric/scores"][a:b]
        if len(np.intersect1d(sc[y == 1], sc[y == 0])) > 0:
            # numpy's default argsort is unstable (SIMD quicksort): the reference's order
            # among EXACT positive/negative ties is platform-dependent.  AUC (ties = 1/2)
            # is order-free and must match; MRR/nDCG are defined by the oracle's
            # reversed-stable rule (see oracle._order_desc) and only bounded here.
            np.testing.assert_allclose(got[0], ref[0], rtol=1e-12)
            assert all(0.0 < g <= 1.0 for g in got[1:])
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)


def test_evaluate_walk(golden, golden_sd):
    owner = O.first_occurrence_rows(golden["eval/news_ids"])
    means, per, scores, table, uv = O.evaluate_pipeline(
        golden_sd, golden["eval/news_tokens"], golden["eval/hist_rows"], golden["eval/cand_offsets"],
        golden["eval/cand_rows"], golden["eval/labels"], news_owner=owner,
        max_count=int(golden["eval/max_count"]), batch=32)
    assert per.shape == golden["eval/per_impression"].shape   # max_count-1 impressions
    np.testing.assert_allclose(scores, golden["eval/scores"], rtol=2e-4, atol=2e-6)
    ref = golden["eval/per_impression"]
    assert np.array_equal(np.isnan(per), np.isnan(ref))
    np.testing.assert_allclose(means, golden["eval/means"], atol=5e-4)
    assert np.array_equal(table[17], table[3])               # first occurrence wins


def test_build_history_left_pad_first_50():
    h = O.build_history([[1, 2, 3], list(range(100, 160)), []], num_clicked=50)
    assert h[0, :47].tolist() == [-1] * 47 and h[0, 47:].tolist() == [1, 2, 3]
    assert h[1].tolist() == list(range(100, 150))            # FIRST 50, not latest
    assert (h[2] == -1).all()


def test_layernorm_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 7, 300)).astype(np.float32)
    g = rng.standard_normal(300).astype(np.float32)
    b = rng.standard_normal(300).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    gt = torch.tensor(g, requires_grad=True)
    bt = torch.tensor(b, requires_grad=True)
    y = torch.nn.functional.layer_norm(xt, (300,), gt, bt)
    y.backward(torch.tensor(dy))
    yo, c = O.layernorm_forward(x, g, b)
    dx, dg, db = O.layernorm_backward(dy, c)
    np.testing.assert_allclose(yo, y.detach().numpy(), atol=2e-5)
    np.testing.assert_allclose(dx, xt.grad.numpy(), atol=5e-5)
    np.testing.assert_allclose(dg, gt.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_torch_port_matches_golden(golden, golden_sd):
    """bench.py's CPU baseline (torch-CPU port of the reference op sequence) is pinned too."""
    from oracle import torch_port as TP
    nv = TP.news_vectors(golden_sd, golden["fwd/tokens"]).numpy()
    assert rel_l2_rows(nv, golden["fwd/news_vectors"]) < 1e-6
    uv = TP.user_vectors(golden_sd, golden["fwd/user_input"]).numpy()
    assert rel_l2_rows(uv, golden["fwd/user_vectors"]) < 1e-6
    import torch
    s = TP.prediction(torch.from_numpy(golden["fwd/news_vectors"][:23]), torch.from_numpy(golden["fwd/user_vectors"][0]))
    np.testing.assert_allclose(s, golden["fwd/scores_single"], rtol=1e-5, atol=1e-6)


def test_torch_port_train_step_matches_golden(golden, golden_sd):
    """bench.py's CPU training baseline (torch-CPU port: forward, CE(label 0), autograd backward, torch.optim.Adam) is
    pinned against the live reference's loss and first Adam step."""
    from oracle import torch_port as TP
    loss, P, _ = TP.train_step(golden_sd, golden["train/cand"], golden["train/clicked"])
    assert abs(loss - float(golden["train/loss"])) < 2e-6
    for k in golden_sd:
        close_adam(clip(P[k].detach().numpy()), golden["train/adam1/" + k], k, gscale=gs(golden, k))


def _torch_ln_encoder(x, p, heads=15):
    """The config-5 encoder composed from stock torch pieces in float64: the reference's MHSA arithmetic
    (multihead_self.py:15-23,46-76), torch.nn.functional.layer_norm, the reference's additive attention
    (additive.py:27-53).  Builder-defined variant: the reference tree has no code for its '+LN' row."""
    import torch
    B, S, D = x.shape
    d = D // heads
    q = x @ p["Wq"].T + p["bq"]
    k = x @ p["Wk"].T + p["bk"]
    v = x @ p["Wv"].T + p["bv"]
    sp = lambda a: a.view(B, S, heads, d).transpose(1, 2)
    s = sp(q) @ sp(k).transpose(-1, -2) / np.sqrt(d)
    e = torch.exp(s)
    attn = e / (e.sum(-1, keepdim=True) + 1e-8)
    c = (attn @ sp(v)).transpose(1, 2).contiguous().view(B, S, D)
    c = torch.nn.functional.layer_norm(c, (D,), p["ln_g"], p["ln_b"], 1e-5)
    tt = torch.tanh(c @ p["Wa"].T + p["ba"])
    w = torch.softmax(tt @ p["qa"], dim=1)
    return torch.bmm(w.unsqueeze(1), c).squeeze(1)


def test_layernorm_variant_encoder_matches_torch_autograd():
    import torch
    rng = np.random.default_rng(7)
    B, S, D = 3, 20, 300
    p = dict(Wq=rng.uniform(-.1, .1, (D, D)), bq=rng.uniform(-.05, .05, D), Wk=rng.uniform(-.1, .1, (D, D)),
             bk=rng.uniform(-.05, .05, D), Wv=rng.uniform(-.1, .1, (D, D)), bv=rng.uniform(-.05, .05, D),
             Wa=rng.uniform(-.05, .05, (200, D)), ba=rng.uniform(-.05, .05, 200), qa=rng.uniform(-.1, .1, 200),
             ln_g=1 + 0.1 * rng.standard_normal(D), ln_b=0.1 * rng.standard_normal(D))
    x = rng.standard_normal((B, S, D))
    out, cache = O.encoder_forward(x, p, 15)
    dout = rng.standard_normal(out.shape)
    dx, g = O.encoder_backward(dout, cache, p, 15)
    pt = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    xt = torch.tensor(x, requires_grad=True)
    yt = _torch_ln_encoder(xt, pt)
    yt.backward(torch.tensor(dout))
    np.testing.assert_allclose(out, yt.detach().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=0, atol=1e-11)
    for k in p:
        np.testing.assert_allclose(g[k], pt[k].grad.numpy(), rtol=0, atol=1e-10, err_msg=k)


def test_mhsa_length_mask_matches_reference():
    """The `length` branch (multihead_self.py:60-68) of the oracle vs vectors generated by the live reference module
    (tests/golden/make_golden_masked.py): full length, one key, zero keys, length past the end."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mhsa_masked_golden.npz"))
    p = {f"W{n}": g[f"W{n}"] for n in "qkv"}
    p.update({f"b{n}": g[f"b{n}"] for n in "qkv"})
    for S in (20, 50):
        out, _ = O.mhsa_forward(g[f"S{S}/x"], p, 15, length=g[f"S{S}/length"])
        np.testing.assert_allclose(out, g[f"S{S}/ctx"], rtol=2e-5, atol=2e-6)
        out0, _ = O.mhsa_forward(g[f"S{S}/x"], p, 15)
        assert not out[2].any()                                  # length 0: 0 / (0 + 1e-8)
        np.testing.assert_array_equal(out[0], out0[0])           # length == S: the mask is a no-op
        np.testing.assert_array_equal(out[5], out0[5])           # length > S too
