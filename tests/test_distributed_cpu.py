"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: news-row sharding + all-gather of the
table, impression sharding balanced by candidates, all-reduce of the metric sums, and the
data-parallel gradient definition (mean of per-rank means == reference step on the concatenated batch).
The arithmetic inside each rank is supplied by the oracle (tests may use it); the product's
sharding / collective code in newsrecommendationsystem_b200.evaluate is what is under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _StubUserEncoder:
    config = None
    precision = "fp32"          # evaluate_tensors then scores from the fp32 table (ops.score_csr, stubbed below)

    def __init__(self, sd):
        self.sd = sd

    def forward_indexed(self, table, rows):
        from oracle import torch_port as TP
        return TP.user_vectors(self.sd, table[rows.long()])


class _StubModel:
    """CPU stand-in with the NRMS inference API; arithmetic from the oracle's torch port."""
    training = False

    def __init__(self, sd):
        self.sd = sd
        self.user_encoder = _StubUserEncoder(sd)

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def get_news_vector(self, news):
        from oracle import torch_port as TP
        return TP.news_vectors(self.sd, news["title"])


def _cpu_score_csr(table, cand_rows, offsets, user_vec):
    imp = torch.repeat_interleave(torch.arange(offsets.numel() - 1), offsets[1:] - offsets[:-1])
    return (table[cand_rows.long()] * user_vec[imp]).sum(-1)


def _cpu_rank_metrics(scores, labels, offsets):
    from oracle import nrms_oracle as O
    n = offsets.numel() - 1
    per = np.zeros((n, 4))
    for i in range(n):
        a, b = int(offsets[i]), int(offsets[i + 1])
        per[i] = O.single_user_metric(labels[a:b].numpy().astype(np.int64), scores[a:b].tolist())
    ok = ~np.isnan(per)
    sums = np.concatenate([np.where(ok, per, 0).sum(0), ok.sum(0).astype(np.float64)])
    return torch.from_numpy(per), torch.from_numpy(sums)


def _eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from newsrecommendationsystem_b200 import evaluate as E, ops, synthetic
        ops.score_csr = _cpu_score_csr
        ops.rank_metrics = _cpu_rank_metrics
        sd = synthetic.init_state_dict(num_words=301, seed=1)
        ntok = synthetic.make_news(61, num_words=301, seed=2)          # 61 rows: uneven shards (31 + 30)
        imp = synthetic.make_impressions(23, 61, seed=3, single_class_every=5, max_cand=25)
        inp = E.EvalInputs(ntok, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"], device="cpu")
        means, det = E.evaluate_tensors(_StubModel(sd), inp, return_details=True)
        # inputs that carry only this rank's block of impressions (what bench.py's e2e leg copies) give the same answer
        host = E.EvalHost(ntok, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
        inp2 = E.EvalInputs.from_host(host, "cpu", shard=True)
        lo, hi = det["impression_range"]
        assert inp2.shard[:2] == (lo, hi) and inp2.hist_rows.shape[0] == hi - lo and inp2.h2d_bytes < host.nbytes()
        means2 = E.evaluate_tensors(_StubModel(sd), inp2)
        np.testing.assert_allclose(means2, means, rtol=1e-12)
        q.put((rank, means, det["table"].numpy().copy(), det["impression_range"]))
    finally:
        dist.destroy_process_group()


class _StubUserEncoder16(_StubUserEncoder):
    precision = "tf32"          # evaluate_tensors then keeps ONE fp16 copy of the news vectors (encode_news_table16)
    layer_norm = None

    def forward_indexed(self, table16, rows):
        from oracle import torch_port as TP
        assert table16.dtype == torch.float16 and table16.shape[1] == 320
        return TP.user_vectors(self.sd, table16[:, :300].float()[rows.long()])


def _cpu_pack_rows_f16(table, out=None):
    n = table.shape[0]
    if out is None:
        out = torch.empty((n + 1, 320), dtype=torch.float16)
    out[:n + 1].zero_()
    out[:n, :300] = table.to(torch.float16)
    out[:n, 300] = 1.0
    return out


def _eval_worker_f16(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from newsrecommendationsystem_b200 import evaluate as E, ops, synthetic
        ops.pack_rows_f16 = _cpu_pack_rows_f16
        ops.score_csr_f16 = lambda t16, cand, offs, uv: _cpu_score_csr(t16[:, :300].float(), cand, offs, uv)
        ops.rank_metrics = _cpu_rank_metrics
        sd = synthetic.init_state_dict(num_words=301, seed=1)
        ntok = synthetic.make_news(61, num_words=301, seed=2)
        imp = synthetic.make_impressions(23, 61, seed=3, single_class_every=5, max_cand=25)
        model = _StubModel(sd)
        model.user_encoder = _StubUserEncoder16(sd)
        host = E.EvalHost(ntok, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
        inp = E.EvalInputs.from_host(host, "cpu", shard=True)          # token rows and impressions of this rank only
        t16, t32 = E.encode_news_table16(model, inp.news_tokens, inp.n_news, inp.news_shard, want_fp32=True)
        means = E.evaluate_tensors(model, inp)
        q.put((rank, means, t16.float().numpy().copy(), t32.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_evaluate_fp16_table_flow_world2():
    """Tensor-mode evaluate across two ranks: every rank packs ITS news vectors to fp16 straight into its slot of the
    gather buffer, one all-gather of 640-byte rows, the same copy feeds the user encoder and the scoring."""
    from newsrecommendationsystem_b200 import synthetic
    from oracle import nrms_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker_f16, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = synthetic.init_state_dict(num_words=301, seed=1)
    ntok = synthetic.make_news(61, num_words=301, seed=2)
    imp = synthetic.make_impressions(23, 61, seed=3, single_class_every=5, max_cand=25)
    ref_means, _, _, ref_table, _ = O.evaluate_pipeline(sd, ntok, imp["hist_rows"], imp["cand_offsets"],
                                                         imp["cand_rows"], imp["labels"])
    (r0, m0, a0, f0), (r1, m1, a1, f1) = res
    from newsrecommendationsystem_b200.evaluate import PAD_REPLICAS as R
    assert a0.shape == (61 + R + 1, 320)                                  # 61 news + R x PADDED_NEWS + the closing row
    assert np.array_equal(a0, a1) and np.array_equal(f0, f1)              # every rank holds the same tables
    np.testing.assert_allclose(a0[:61, :300], ref_table[:61], rtol=2e-3, atol=2e-4)     # fp16 rounding of the rows
    assert np.all(a0[:61 + R, 300] == 1.0) and not a0[:61, 301:].any() and not a0[61:61 + R, :300].any() and not a0[61 + R].any()
    np.testing.assert_allclose(f0[:61], ref_table[:61], rtol=2e-4, atol=2e-6)
    assert not f0[61].any()
    np.testing.assert_allclose(m0, m1, rtol=1e-12)
    np.testing.assert_allclose(m0, ref_means, atol=5e-3)


def test_evaluate_sharding_allgather_allreduce_world2():
    from newsrecommendationsystem_b200 import synthetic
    from oracle import nrms_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = synthetic.init_state_dict(num_words=301, seed=1)
    ntok = synthetic.make_news(61, num_words=301, seed=2)
    imp = synthetic.make_impressions(23, 61, seed=3, single_class_every=5, max_cand=25)
    ref_means, _, _, ref_table, _ = O.evaluate_pipeline(sd, ntok, imp["hist_rows"], imp["cand_offsets"],
                                                         imp["cand_rows"], imp["labels"])
    (r0, m0, t0, range0), (r1, m1, t1, range1) = res
    np.testing.assert_allclose(t0, t1)                                    # every rank holds the full table
    np.testing.assert_allclose(t0[:61], ref_table[:61], rtol=2e-4, atol=2e-6)
    assert not t0[61].any()                                               # PADDED_NEWS row stays zero
    assert range0[0] == 0 and range0[1] == range1[0] and range1[1] == 23  # contiguous impression blocks
    np.testing.assert_allclose(m0, m1, rtol=1e-12)                        # all-reduced metric means agree
    np.testing.assert_allclose(m0, ref_means, atol=5e-4)


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from newsrecommendationsystem_b200 import synthetic
        from oracle import nrms_oracle as O
        sd = synthetic.init_state_dict(num_words=201, seed=4)
        ntok = synthetic.make_news(40, num_words=201, seed=5)
        cand, clicked = synthetic.make_train_batch(4, ntok, k_neg=4, seed=6)
        lo, hi = rank * 2, rank * 2 + 2                                   # equal shards of the global batch
        logits, cache = O.nrms_forward(sd, cand[lo:hi], clicked[lo:hi])
        loss, dlog = O.cross_entropy_label0(logits)
        grads = O.nrms_backward(dlog, cache, sd)
        key = "user_encoder.multihead_self_attention.W_V.weight"
        g = torch.from_numpy(grads[key].copy())
        dist.all_reduce(g)                                                # what FusedAdam.allreduce_grads does
        g *= 1.0 / world                                                  # the scale folded into nrms_adam_step
        q.put((rank, g.numpy(), float(loss)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_is_reference_step_on_concatenated_batch():
    from newsrecommendationsystem_b200 import synthetic
    from oracle import nrms_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = synthetic.init_state_dict(num_words=201, seed=4)
    ntok = synthetic.make_news(40, num_words=201, seed=5)
    cand, clicked = synthetic.make_train_batch(4, ntok, k_neg=4, seed=6)
    logits, cache = O.nrms_forward(sd, cand, clicked)
    loss, dlog = O.cross_entropy_label0(logits)
    grads = O.nrms_backward(dlog, cache, sd)
    key = "user_encoder.multihead_self_attention.W_V.weight"
    np.testing.assert_allclose(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][1], grads[key], rtol=2e-4, atol=1e-7)
    assert abs((res[0][2] + res[1][2]) / 2 - float(loss)) < 1e-6
