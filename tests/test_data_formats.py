"""newsrecommendationsystem_b200/data.py against what the LIVE reference dataset classes make of the same files
(tests/golden/make_golden_data.py: dataset.BaseDataset, evaluate.NewsDataset, evaluate.BehaviorsDataset): integer and
string work, so everything is compared exactly."""
import os

import numpy as np
import pytest

from conftest import ROOT

DATA = os.path.join(ROOT, "tests", "golden", "data")


@pytest.fixture(scope="module")
def gd():
    z = np.load(os.path.join(ROOT, "tests", "golden", "data_golden.npz"))
    return {k: z[k] for k in z.files}


def test_news_table_matches_reference_news_dataset(gd):
    from newsrecommendationsystem_b200.data import load_news_parsed
    news = load_news_parsed(os.path.join(DATA, "news_parsed.tsv"), ("title", "abstract", "category", "subcategory"))
    assert news.ids == [str(x) for x in gd["news/ids"]]
    assert np.array_equal(news.title, gd["news/titles"]) and news.title.dtype == np.int64
    assert news.columns["abstract"].shape == (len(news), 50) and news.columns["category"].shape == (len(news),)
    # the repeated id maps to its FIRST row (evaluate.py:197-201)
    assert news.ids[9] == news.ids[2] and news.row_of[news.ids[9]] == 2
    t = news.token_table_with_pad()
    assert t.shape == (len(news) + 1, 20) and not t[-1].any()
    with pytest.raises(ValueError):
        load_news_parsed(os.path.join(DATA, "news_parsed.tsv"), ("title_entities",))


def test_training_rows_reproduce_the_reference_minibatch_tensors(gd):
    """cand_rows / hist_rows index the token table to exactly the title tensors BaseDataset.__getitem__ stacks
    (first 50 clicks, left padding with the all-zero title, dataset.py:62-85)."""
    from newsrecommendationsystem_b200.data import load_news_parsed, load_behaviors_parsed
    news = load_news_parsed(os.path.join(DATA, "news_parsed.tsv"))
    tr = load_behaviors_parsed(os.path.join(DATA, "behaviors_parsed.tsv"), news)
    table = news.token_table_with_pad()
    assert tr.cand_rows.shape == (7, 3) and tr.hist_rows.shape == (7, 50)
    assert np.array_equal(table[tr.cand_rows], gd["train/cand_titles"])
    assert np.array_equal(table[tr.hist_rows], gd["train/clicked_titles"])
    assert np.array_equal(tr.clicked, np.tile(np.array([1, 0, 0], dtype=np.int8), (7, 1)))
    assert (tr.hist_rows[0, :47] == len(news)).all() and (tr.hist_rows[2] < len(news)).all()      # 3 clicks / 64 clicks


def test_evaluate_rows_match_reference_behaviors(gd):
    from newsrecommendationsystem_b200.data import load_news_parsed, load_behaviors
    from newsrecommendationsystem_b200.evaluate import EvalHost
    news = load_news_parsed(os.path.join(DATA, "news_parsed.tsv"))
    ev = load_behaviors(os.path.join(DATA, "behaviors.tsv"), news)
    for k in ("hist_rows", "cand_offsets", "cand_rows", "labels"):
        assert np.array_equal(getattr(ev, k), gd["eval/" + k]), k
    assert ev.clicked_news_strings == [str(x) for x in gd["eval/keys"]]
    assert (ev.hist_rows[0] == -1).all()                                    # empty history -> 50 x PADDED_NEWS
    assert (ev.hist_rows[3] >= 0).all()                                     # 77 clicks -> the FIRST 50
    # max_count: the reference breaks BEFORE the max_count-th impression (evaluate.py:247-249)
    assert len(load_behaviors(os.path.join(DATA, "behaviors.tsv"), news, max_count=4).impression_ids) == 3
    # and the tables feed the device-resident evaluate as they are
    host = EvalHost(news.title, ev.hist_rows, ev.cand_offsets, ev.cand_rows, ev.labels, news_ids=news.ids)
    assert host.n_impressions == 6 and host.n_news == len(news)


def test_evaluate_directory_entry_point_on_cpu_stub(gd, monkeypatch):
    """evaluate(model, directory, num_workers, max_count) -- the reference's signature (evaluate.py:171-272) -- over the sample
    files: file reading, first-wins ids, pad spreading and the pipeline plumbing, with the arithmetic of every stage supplied
    by the oracle's torch port (CPU stand-ins of the library calls, as in test_distributed_cpu.py)."""
    import sys
    import torch
    from newsrecommendationsystem_b200 import evaluate as E, ops, synthetic
    from oracle import nrms_oracle as O
    from test_distributed_cpu import _StubModel, _cpu_score_csr, _cpu_rank_metrics
    monkeypatch.setattr(ops, "score_csr", _cpu_score_csr)
    monkeypatch.setattr(ops, "rank_metrics", _cpu_rank_metrics)
    sd = synthetic.init_state_dict(num_words=401, seed=1)
    model = _StubModel(sd)
    model.parameters = lambda: iter([torch.zeros(1)])
    means = E.evaluate(model, DATA, num_workers=4, max_count=sys.maxsize)
    # (the fixture's history / candidate rows already name the FIRST row of a repeated id)
    ref_means = O.evaluate_pipeline(sd, gd["news/titles"], gd["eval/hist_rows"], gd["eval/cand_offsets"], gd["eval/cand_rows"],
                                    gd["eval/labels"])[0]
    np.testing.assert_allclose(means, ref_means, atol=1e-6)
    means3 = E.evaluate(model, DATA, 0, max_count=4)                      # impressions 1..3 only
    ref3 = O.evaluate_pipeline(sd, gd["news/titles"], gd["eval/hist_rows"], gd["eval/cand_offsets"], gd["eval/cand_rows"],
                               gd["eval/labels"], max_count=4)[0]
    np.testing.assert_allclose(means3, ref3, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,atol", [("fp32", 1e-5), ("tf32", 2e-3)])
def test_evaluate_directory_entry_point_on_gpu(gd, precision, atol):
    """The same call on the CUDA path (both precision modes) against the oracle's evaluate walk."""
    import torch
    from newsrecommendationsystem_b200 import NRMS, NRMSConfig, evaluate as E, synthetic
    from oracle import nrms_oracle as O

    class Cfg(NRMSConfig):
        num_words = 401
    sd = synthetic.init_state_dict(num_words=401, seed=1)
    m = NRMS(Cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.to("cuda:0").eval().set_precision(precision)
    means = E.evaluate(m, DATA, num_workers=4)
    ref = O.evaluate_pipeline(sd, gd["news/titles"], gd["eval/hist_rows"], gd["eval/cand_offsets"], gd["eval/cand_rows"],
                              gd["eval/labels"])[0]
    np.testing.assert_allclose(means, ref, atol=atol)


def test_load_target_user_takes_the_first_row(tmp_path):
    from newsrecommendationsystem_b200.recommend import load_target_user
    f = tmp_path / "behaviors.tsv"
    f.write_text("1\tU1\tt\tN1 N2\tN3-0 N4-1\n2\tU2\tt\t\tN5-0\n3\tU1\tt\tN9\tN8-1\n")
    assert load_target_user(str(f), "U1") == (["N1", "N2"], ["N3-0", "N4-1"])
    assert load_target_user(str(f), "U2") == ([], ["N5-0"])
    with pytest.raises(KeyError):
        load_target_user(str(f), "U7")


@pytest.mark.gpu
def test_recommender_from_directory_and_cache_roundtrip(gd, tmp_path):
    """Recommender.from_directory on the sample files: encodes the news, writes the reference's news2vector.pt, reloads it,
    and ranks the first impression of a user like the oracle does."""
    import shutil
    import torch
    from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic
    from newsrecommendationsystem_b200.recommend import Recommender
    from oracle import nrms_oracle as O

    class Cfg(NRMSConfig):
        num_words = 401
    for name in ("news_parsed.tsv", "behaviors.tsv"):
        shutil.copy(os.path.join(DATA, name), tmp_path / name)
    sd = synthetic.init_state_dict(num_words=401, seed=1)
    m = NRMS(Cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.to("cuda:0").eval().set_precision("fp32")
    rec = Recommender.from_directory(m, str(tmp_path))
    cache = torch.load(tmp_path / "news2vector.pt", weights_only=False)
    assert "PADDED_NEWS" in cache and not cache["PADDED_NEWS"].any() and len(cache) == len(set(gd["news/ids"].tolist())) + 1
    rec2 = Recommender.from_directory(m, str(tmp_path))                     # now from the cache file
    table = O.evaluate_pipeline(sd, gd["news/titles"], gd["eval/hist_rows"][:0], gd["eval/cand_offsets"][:1],
                                gd["eval/cand_rows"][:0], gd["eval/labels"][:0])[3]
    for r in (rec, rec2):
        ids, y = r.recommend_target_user(str(tmp_path), "U3")              # impression 4: 77 clicks -> the first 50
        a, b = int(gd["eval/cand_offsets"][3]), int(gd["eval/cand_offsets"][4])
        _, y_ref, order_ref = O.recommend_user(sd, table, np.where(gd["eval/hist_rows"][3] < 0, len(table) - 1,
                                                                   gd["eval/hist_rows"][3]), gd["eval/cand_rows"][a:b])
        np.testing.assert_allclose(y, y_ref[order_ref], atol=1e-5)
        assert len(ids) == b - a

