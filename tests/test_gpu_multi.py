"""NCCL world-size-2 tests of the PRODUCT's multi-GPU paths (run with `gpurun --gpus 2 -- pytest -m gpu
tests/test_gpu_multi.py`; skipped on a one-GPU box, where bench.py's `metrics_match_1rank` carries the same check):

  * evaluate: news rows sharded + NCCL all-gather of the fp16 table, impressions sharded by candidate
    count, 8-double all-reduce -- the metric means at 2 ranks equal the 1-rank means of the same library and, through
    those, the oracle's evaluate walk (tests/test_gpu_parity_full.py pins the 1-rank means against the reference port);
  * training: two ranks x half a batch through TrainStep (flat-gradient all-reduce, 1/world inside the
    Adam kernel) == one rank on the concatenated batch, dropout off (the reference contract: CE mean of equal shards).
"""
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

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240),
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                 reason="needs two GPUs (gpurun --gpus 2)")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Cfg:
    num_words = 2001
    word_embedding_dim = 300
    num_attention_heads = 15
    query_vector_dim = 200
    dropout_probability = 0.0
    num_words_title = 20
    num_clicked_news_a_user = 50


def _model(dev, train=False):
    from newsrecommendationsystem_b200 import NRMS, synthetic
    sd = synthetic.init_state_dict(num_words=_Cfg.num_words, seed=3)
    m = NRMS(_Cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.to(dev).set_precision("tf32")
    return m.train() if train else m.eval()


def _data():
    from newsrecommendationsystem_b200 import synthetic
    news = synthetic.make_news(6000, num_words=_Cfg.num_words, seed=5)
    imp = synthetic.make_impressions(4000, 6000, seed=6)
    return news, imp


def _eval_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
        dev = torch.device("cuda", rank)
        news, imp = _data()
        host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
        means = evaluate_tensors(_model(dev), EvalInputs.from_host(host, dev, shard=True))
        if rank == 0:
            np.save(out_path, np.asarray([float(x) for x in means]))
    finally:
        dist.destroy_process_group()


def test_evaluate_two_ranks_equals_one_rank(tmp_path):
    from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
    dev = torch.device("cuda", 0)
    news, imp = _data()
    host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
    one = [float(x) for x in evaluate_tensors(_model(dev), EvalInputs.from_host(host, dev))]
    out_path = str(tmp_path / "means.npy")          # results travel through files: a Queue.put of a large array blocks the
    mp.spawn(_eval_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)      # child until the parent reads
    two = np.load(out_path)
    np.testing.assert_allclose(two, one, rtol=0, atol=1e-9)      # same kernels on the same rows: only the fp64 sum order differs


def _train_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from newsrecommendationsystem_b200.train import TrainStep
        dev = torch.device("cuda", rank)
        titles = _train_batch()
        half = titles.shape[0] // world
        ts = TrainStep(_model(dev, train=True), lr=1e-4)
        ts.step_tokens(torch.from_numpy(titles[rank * half:(rank + 1) * half]), 5)
        torch.cuda.synchronize()
        if rank == 0:
            # after the step flat_grad holds the all-reduced SUM of the ranks' gradients; the Adam kernel applied 1/world
            np.savez(out_path, g=ts.optimizer.flat_grad.detach().cpu().numpy() / world,
                     p=ts.optimizer.flat_param.detach().cpu().numpy())
    finally:
        dist.destroy_process_group()


def _train_batch():
    rng = np.random.default_rng(12)
    t = rng.integers(1, _Cfg.num_words, size=(32, 55, 20)).astype(np.int64)
    t[:, :, 12:] = 0
    return t


def test_data_parallel_step_equals_full_batch_step(tmp_path):
    from newsrecommendationsystem_b200.train import TrainStep
    dev = torch.device("cuda", 0)
    ts = TrainStep(_model(dev, train=True), lr=1e-4)
    ts.step_tokens(torch.from_numpy(_train_batch()), 5)
    torch.cuda.synchronize()
    g1, p1 = ts.optimizer.flat_grad.detach().cpu().numpy(), ts.optimizer.flat_param.detach().cpu().numpy()
    out_path = str(tmp_path / "dp.npz")
    mp.spawn(_train_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    z = np.load(out_path)
    g2, p2 = z["g"], z["p"]
    # averaged two-rank gradient == full-batch gradient (TF32 contractions and fp32 atomics: up to summation order)
    scale = float(np.abs(g1).max())
    assert float(np.abs(g1 - g2).max()) <= 2e-4 * scale, float(np.abs(g1 - g2).max()) / scale
    # and the parameters moved the same way wherever the gradient is not rounding noise (the first Adam step is
    # lr * sign(g): an entry whose sign is noise may differ by 2 lr)
    solid = np.abs(g1) > 1e-4 * scale
    assert float(np.abs(p1 - p2)[solid].max()) <= 2e-6
    assert float(np.abs(p1 - p2).max()) <= 2.1e-4
