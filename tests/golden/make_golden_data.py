"""Fixtures for newsrecommendationsystem_b200/data.py: tiny files in the reference's on-disk formats plus what the LIVE
reference dataset classes make of them (run in the build container only: needs /root/reference).

    python tests/golden/make_golden_data.py   ->  tests/golden/data/{news_parsed,behaviors_parsed,behaviors}.tsv, data_golden.npz

Pinned: `dataset.BaseDataset.__getitem__` (src/dataset.py:62-85: candidate / clicked title tensors of every training sample,
first 50 clicks, left padding), `evaluate.NewsDataset` (src/evaluate.py:51-78) and `evaluate.BehaviorsDataset` (:127-157).
`evaluate.UserDataset` raises under this container's pandas; its history rule (`clicked_news.split()[:50]`,
left-padded with 'PADDED_NEWS', :117-124) is applied here to BehaviorsDataset's `clicked_news_string`.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")

from dataset import BaseDataset  # noqa: E402  (reference; MODEL_NAME defaults to NRMS)
import evaluate as ref_eval  # noqa: E402  (reference)


def main():
    os.makedirs(DATA, exist_ok=True)
    rng = np.random.default_rng(3)
    n_news = 14
    ids = [f"N{100 + i}" for i in range(n_news)]
    ids[9] = ids[2]                       # a repeated id: the first row wins (evaluate.py:197-201)
    with open(os.path.join(DATA, "news_parsed.tsv"), "w") as f:
        f.write("id\tcategory\tsubcategory\ttitle\tabstract\ttitle_entities\tabstract_entities\n")
        for i, nid in enumerate(ids):
            ln = int(rng.integers(1, 21))
            title = list(map(int, rng.integers(1, 400, ln))) + [0] * (20 - ln)
            abstract = list(map(int, rng.integers(1, 400, 50)))
            f.write(f"{nid}\t{int(rng.integers(1, 9))}\t{int(rng.integers(9, 30))}\t{title}\t{abstract}\t{[0] * 20}\t{[0] * 50}\n")
    uniq = [x for i, x in enumerate(ids) if x not in ids[:i]]
    with open(os.path.join(DATA, "behaviors_parsed.tsv"), "w") as f:
        f.write("user\tclicked_news\tcandidate_news\tclicked\n")
        for b in range(7):
            n_click = [3, 50, 64, 1, 17, 50, 8][b]
            clicks = " ".join(rng.choice(uniq, n_click))
            cand = " ".join(rng.choice(uniq, 3))
            f.write(f"{b + 1}\t{clicks}\t{cand}\t1 0 0\n")
    with open(os.path.join(DATA, "behaviors.tsv"), "w") as f:
        for i in range(6):
            n_click = [0, 5, 50, 77, 1, 12][i]
            clicks = " ".join(rng.choice(uniq, n_click))
            imps = " ".join(f"{x}-{int(rng.random() < 0.3)}" for x in rng.choice(uniq, int(rng.integers(2, 9))))
            f.write(f"{i + 1}\tU{i}\t11/1{i}/2019 9:0{i}:00 AM\t{clicks}\t{imps}\n")

    out = {}
    # ---- training samples through the reference's BaseDataset --------------------------------------------
    # (its news frame is indexed by id: a repeated id makes to_dict('index') raise, so the training fixture reads a
    #  de-duplicated copy -- first occurrence kept, as evaluate does)
    dedup = os.path.join(DATA, "_news_dedup.tsv")
    seen = set()
    with open(os.path.join(DATA, "news_parsed.tsv")) as fi, open(dedup, "w") as fo:
        for k, line in enumerate(fi):
            nid = line.split("\t", 1)[0]
            if k == 0 or nid not in seen:
                fo.write(line)
            seen.add(nid)
    ds = BaseDataset(os.path.join(DATA, "behaviors_parsed.tsv"), dedup)
    cand, clicked = [], []
    for i in range(len(ds)):
        item = ds[i]
        cand.append(torch.stack([x["title"] for x in item["candidate_news"]]).numpy())
        clicked.append(torch.stack([x["title"] for x in item["clicked_news"]]).numpy())
    out["train/cand_titles"] = np.stack(cand)
    out["train/clicked_titles"] = np.stack(clicked)
    os.unlink(dedup)
    # ---- evaluate-side datasets ------------------------------------------------------------------------------
    nd = ref_eval.NewsDataset(os.path.join(DATA, "news_parsed.tsv"))
    out["news/ids"] = np.array([nd[i]["id"] for i in range(len(nd))])
    out["news/titles"] = np.stack([nd[i]["title"].numpy() for i in range(len(nd))])
    bd = ref_eval.BehaviorsDataset(os.path.join(DATA, "behaviors.tsv"))
    first = {}
    for r, nid in enumerate(out["news/ids"]):
        first.setdefault(str(nid), r)                                   # evaluate.py:197-201
    hist, offs, rows, labels, keys = [], [0], [], [], []
    for i in range(len(bd)):
        item = bd[i]
        key = item["clicked_news_string"]
        if not isinstance(key, str):      # `fillna(' ', inplace=True)` (evaluate.py:142) is a no-op under pandas 3: NaN stays;
            key = " "                     # the value the reference means is ' '
        clicks = key.split()[:50]                                       # UserDataset rule, evaluate.py:117
        hist.append([-1] * (50 - len(clicks)) + [first[x] for x in clicks])
        for news in item["impressions"]:
            rows.append(first[news.split('-')[0]])                      # evaluate.py:252
            labels.append(int(news.split('-')[1]))                      # :261-263
        offs.append(len(rows))
        keys.append(key)
    out["eval/hist_rows"] = np.asarray(hist, dtype=np.int64)
    out["eval/cand_offsets"] = np.asarray(offs, dtype=np.int64)
    out["eval/cand_rows"] = np.asarray(rows, dtype=np.int64)
    out["eval/labels"] = np.asarray(labels, dtype=np.int8)
    out["eval/keys"] = np.array(keys)
    np.savez_compressed(os.path.join(HERE, "data_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "data_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
