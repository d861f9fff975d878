"""CPU-side checks: the C-ABI library loads and exports every symbol include/nrms_b200.h declares
(no compute calls), the ctypes table matches the header, argument validation returns error codes
without touching a GPU, and the host-side evaluate helpers follow the reference's indexing rules."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrms_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nrms_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from newsrecommendationsystem_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from newsrecommendationsystem_b200 import _lib
    declared = header_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nrms_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"


def test_argument_count_matches_header():
    from newsrecommendationsystem_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", src)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_version_and_error_paths_without_gpu(lib):
    assert lib.nrms_abi_version() == 2
    # invalid arguments are rejected before any CUDA call; message is retrievable
    rc = lib.nrms_score_fwd(None, None, 4, 0, 300, None, None)
    assert rc == 1 and b"bad sizes" in lib.nrms_last_error()
    rc = lib.nrms_news_encoder_fwd(None, 8, 21, None, 10, None, None, None, None, None, None, None, None, 0,
                                   0.0, 0, 0, 0, None)
    assert rc == 2 and b"unsupported" in lib.nrms_last_error()
    rc = lib.nrms_adam_step(None, None, None, None, 16, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 0, 1.0, None)
    assert rc == 1
    assert lib.nrms_encoder_stash_bytes(7040, 20) >= 7040 * 20 * (300 + 900 + 300 + 200 + 1) * 4
    assert lib.nrms_encoder_bwd_workspace_bytes(128, 50, 0) > 0


def test_missing_library_fails_loudly(monkeypatch):
    from newsrecommendationsystem_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnrms_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "newsrecommendationsystem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), os.path.join(dirpath, f)


def test_first_occurrence_and_history_match_oracle():
    from newsrecommendationsystem_b200 import evaluate as E
    from oracle import nrms_oracle as O
    ids = np.array([5, 9, 5, 7, 9, 9, 1])
    assert E.first_occurrence_rows(ids).tolist() == O.first_occurrence_rows(ids).tolist() == [0, 1, 0, 3, 1, 1, 6]
    hs = [[1, 2, 3], list(range(100, 160)), []]
    assert np.array_equal(E.build_history(hs), O.build_history(hs))


def test_sharding_helpers():
    from newsrecommendationsystem_b200 import evaluate as E
    assert [E.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [E.shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    offs = np.array([0, 10, 12, 40, 41, 80, 100])
    b = E.shard_impressions_by_candidates(offs, 4)
    assert b[0] == 0 and b[-1] == 6 and (np.diff(b) >= 0).all()
    loads = [offs[b[i + 1]] - offs[b[i]] for i in range(4)]
    assert sum(loads) == 100 and max(loads) <= 60


def test_synthetic_shapes():
    from newsrecommendationsystem_b200 import synthetic
    toks = synthetic.make_news(1000)
    assert toks.shape == (1000, 20) and toks.dtype == np.int64 and (toks[:, 0] > 0).all() and toks.max() < 70976
    imp = synthetic.make_impressions(3000, 1000)
    C_ = np.diff(imp["cand_offsets"])
    assert C_.min() >= 2 and C_.max() <= 300 and 25 < C_.mean() < 50
    h = imp["hist_rows"]
    pads = (h < 0)
    assert (pads[:, :-1] >= pads[:, 1:]).all()          # pads are a LEFT prefix
    lab = imp["labels"]
    o = imp["cand_offsets"]
    assert lab[o[0]:o[1]].max() == 1 and lab[o[999]:o[1000]].max() == 0   # every 1000th single-class
    cand, clicked = synthetic.make_train_batch(16, toks, k_neg=4)
    assert cand.shape == (16, 5, 20) and clicked.shape == (16, 50, 20)


def test_news2vector_cache_round_trip():
    """recommend.py:211-243 keeps a dict id -> vector with a PADDED_NEWS zero entry; the table form keeps that row last."""
    import torch
    from newsrecommendationsystem_b200 import checkpoint as ck
    table = torch.arange(12, dtype=torch.float32).view(4, 3)
    table[3] = 0
    n2v = ck.news2vector_from_table(["N1", "N2", "N1"], table)        # duplicate id: the first row wins
    assert set(n2v) == {"N1", "N2", "PADDED_NEWS"} and torch.equal(n2v["N1"], table[0])
    assert not n2v["PADDED_NEWS"].any()
    ids, back = ck.table_from_news2vector(n2v)
    assert ids == ["N1", "N2"] and back.shape == (3, 3) and torch.equal(back[:2], table[:2]) and not back[2].any()


def test_out_of_range_indices_raise_like_the_reference():
    """The reference raises on unknown ids (KeyError in news2vector[...], evaluate.py:221,252; IndexError inside
    nn.Embedding, news_encoder.py:38).  The device kernels do not check, so the host formats do."""
    from newsrecommendationsystem_b200 import synthetic
    from newsrecommendationsystem_b200.evaluate import EvalHost
    ntok = synthetic.make_news(30, num_words=101, seed=2)
    imp = synthetic.make_impressions(9, 30, seed=3, max_cand=12)
    EvalHost(ntok, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"], num_words=101)      # in range
    bad = imp["hist_rows"].copy(); bad[2, 49] = 30
    with pytest.raises(IndexError):
        EvalHost(ntok, bad, imp["cand_offsets"], imp["cand_rows"], imp["labels"])
    bad = imp["cand_rows"].copy(); bad[5] = -1
    with pytest.raises(IndexError):
        EvalHost(ntok, imp["hist_rows"], imp["cand_offsets"], bad, imp["labels"])
    bad = ntok.copy(); bad[3, 0] = 101
    with pytest.raises(IndexError):
        EvalHost(bad, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"], num_words=101)


def test_checkpoint_keeps_the_dropout_stream_position(tmp_path):
    """save_checkpoint / load_checkpoint carry the Philox dropout counter next to the reference's four keys."""
    import torch
    from newsrecommendationsystem_b200 import checkpoint

    class _Enc:
        dropout_seed, _dropout_calls = 0x5EED, 12345

    class _Model:
        news_encoder = _Enc()

        def state_dict(self):
            return {"w": torch.ones(2)}

        def load_state_dict(self, sd):
            assert set(sd) == {"w"}

    class _Opt:
        def state_dict(self):
            return {"state": {}, "param_groups": []}

    path = str(tmp_path / "ckpt-7.pth")
    checkpoint.save_checkpoint(path, _Model(), _Opt(), 7, 0.5)
    raw = torch.load(path, weights_only=False)
    assert {"model_state_dict", "optimizer_state_dict", "step", "early_stop_value"} <= set(raw)
    m2 = _Model()
    m2.news_encoder = type("E", (), {"dropout_seed": 1, "_dropout_calls": 0})()
    assert checkpoint.load_checkpoint(path, m2) == (7, 0.5)
    assert (m2.news_encoder.dropout_seed, m2.news_encoder._dropout_calls) == (0x5EED, 12345)


def test_pad_references_are_spread_over_the_replica_rows():
    """PADDED_NEWS is PAD_REPLICAS identical zero rows behind the news rows; EvalHost sends every pad (-1) to one of them
    (an L2 hot spot otherwise) and leaves real rows alone."""
    from newsrecommendationsystem_b200 import evaluate as E
    rng = np.random.default_rng(0)
    n_news = 37
    hist = rng.integers(0, n_news, size=(200, 50)).astype(np.int64)
    hist[np.arange(50)[None, :] < rng.integers(0, 51, size=200)[:, None]] = -1
    ref = hist.copy()
    tok = rng.integers(1, 90, size=(n_news, 20)).astype(np.int64)
    offs = np.arange(0, 401, 2, dtype=np.int64)
    host = E.EvalHost(tok, hist, offs, rng.integers(0, n_news, size=400), np.zeros(400, dtype=np.int8))
    h = host.hist_rows.numpy()
    assert np.array_equal(h[ref >= 0], ref[ref >= 0])
    pads = h[ref < 0]
    assert pads.min() >= n_news and pads.max() < n_news + E.PAD_REPLICAS
    assert len(np.unique(pads)) == E.PAD_REPLICAS                       # all replicas are in use
    assert np.bincount(pads - n_news).max() < 3 * len(pads) / E.PAD_REPLICAS     # and evenly
    assert np.array_equal(hist, ref)                                    # the caller's array is not modified


def test_gemm_work_mapping_covers_every_tile_once():
    """Restatement of the work mapping of csrc/tc_gemm.cu (host grid rule + the kernel's item_w): a CTA takes row blocks
    (m tile, K split) and walks all n tiles of each -- or, under half a wave of row blocks, the n tiles are dealt out singly.
    Every (m tile, n tile, split) must be visited exactly once for every shape the path uses."""
    SM, BM, BN = 148, 128, 256

    def visited(M, N, k_splits):
        m_tiles, n_tiles = (M + BM - 1) // BM, (N + BN - 1) // BN
        items = m_tiles * k_splits
        grid = SM
        if 2 * items < grid:
            items *= n_tiles
        grid = min(grid, items)
        total_rb = m_tiles * k_splits
        ng = n_tiles if total_rb >= grid else 1
        total_groups = total_rb * (n_tiles // ng)
        seen = []
        for b in range(grid):
            my_groups = (total_groups - b + grid - 1) // grid if b < total_groups else 0
            for i in range(my_groups * ng):
                grp = b + (i // ng) * grid
                rb, n = (grp // n_tiles, grp % n_tiles) if ng == 1 else (grp, i % ng)
                seen.append(((rb // k_splits) * n_tiles + n) * k_splits + rb % k_splits)
        return sorted(seen), m_tiles * n_tiles * k_splits

    for M in (1, 61, 128, 129, 1000, 6400, 9472, 65239, 140800, 900, 200):
        for N in (16, 200, 256, 300, 900, 1080):
            for ks in (1, 2, 10, 18, 74):
                seen, total = visited(M, N, ks)
                assert seen == list(range(total)), (M, N, ks)
