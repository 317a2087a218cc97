"""BASELINE config 1: DeepFM k=8, batch 256 on MovieLens-100k (the data set the reference bundles under data/ml-100k).

The loader / feature builder mirror recman/examples/datasets/ml_100k.py and recman/examples/utils.py:30-75.  The GPU box
has no copy of the reference tree, so the tests use tests/golden/ml100k_sample.csv.gz - a seeded 3000-row sample of the
merged ua.base frame written by tests/golden/make_ml100k_sample.py - plus a synthetic directory in the ml-100k file format.
"""
import json
import os

import numpy as np
import pandas as pd
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sample():
    df = pd.read_csv(os.path.join(GOLDEN, "ml100k_sample.csv.gz"), dtype={"zip": str})
    meta = json.load(open(os.path.join(GOLDEN, "ml100k_sample_meta.json")))
    df["label"] = (df.rating >= 4).astype(np.float32)
    return df, meta["domains"]


def _write_ml100k_dir(root):
    d = root / "ml-100k"
    d.mkdir()
    genres = ["unknown", "Action", "Comedy", "Drama"]
    (d / "u.genre").write_text("".join(f"{g}|{i}\n" for i, g in enumerate(genres)))
    (d / "u.occupation").write_text("artist\nengineer\n")
    (d / "u.user").write_text("1|24|M|engineer|85711\n2|53|F|artist|94043\n3|23|M|artist|32067\n")
    items = [(1, "Toy Story (1995)", [0, 0, 1, 0]), (2, "GoldenEye (1995)", [0, 1, 0, 1]), (3, "Nothing (1996)", [0, 0, 0, 0])]
    (d / "u.item").write_text("".join(
        f"{i}|{t}|01-Jan-1995||http://x/{i}|" + "|".join(map(str, fl)) + "\n" for i, t, fl in items))
    (d / "ua.base").write_text("1\t1\t5\t874965758\n1\t2\t3\t876893171\n2\t1\t4\t888550871\n3\t3\t2\t891350008\n")
    (d / "ua.test").write_text("2\t2\t5\t888552084\n")
    return genres


def test_loader_merges_users_items_and_builds_genre_strings(tmp_path):
    from recman_b200.examples.datasets.ml_100k import get_data

    genres = _write_ml100k_dir(tmp_path)
    df_train, df_test, domains = get_data(str(tmp_path))
    assert domains == dict(genres=genres, occupations=["artist", "engineer"])
    assert list(df_train.columns) == ["user_id", "item_id", "rating", "timestamp", "age", "gender", "occupation", "zip",
                                      "title", "release_date", "genres"]
    assert len(df_train) == 4 and len(df_test) == 1
    by = {(r.user_id, r.item_id): r for r in df_train.itertuples()}
    assert by[(1, 1)].genres == "Comedy" and by[(1, 2)].genres == "Action|Drama" and by[(3, 3)].genres == ""
    assert by[(2, 1)].age == 53 and by[(2, 1)].occupation == "artist" and by[(1, 2)].rating == 3
    assert df_test.iloc[0].genres == "Action|Drama" and df_test.iloc[0].gender == "F"


def test_get_ml_dataset_split_and_labels(tmp_path):
    from recman_b200.examples.utils import get_ml_dataset

    _write_ml100k_dir(tmp_path)
    df_train, df_valid, df_test, _ = get_ml_dataset(str(tmp_path), frac=1.0, random_seed=2019)
    assert len(df_train) + len(df_valid) == 4 and len(df_train) == 3 and len(df_test) == 1
    assert set(df_train.index).isdisjoint(df_valid.index)
    both = pd.concat([df_train, df_valid])
    assert ((both.rating >= 4).astype(np.float32) == both.label).all() and df_test.label.iloc[0] == 1.0


def test_create_ml_features_on_the_sample():
    from recman_b200.examples.utils import create_ml_features
    from recman_b200.th.input import DenseFeat, MultiValCsvFeat, SparseFeat

    df, domains = _sample()
    fd = create_ml_features(df, domains)
    assert list(fd) == ["user_id", "item_id", "gender", "occupation", "zip", "timestamp", "age", "genres"]
    for name in ("user_id", "item_id", "gender", "occupation", "zip"):
        assert type(fd[name]) is SparseFeat and fd[name].feat_size == df[name].nunique() + 1  # row 0 = unknown
        ids = fd[name](df[name].values)
        assert ids.shape == (len(df), 1) and ids.min() >= 1 and ids.max() == df[name].nunique()
        assert int(fd[name](np.asarray([df[name].iloc[0]]))[0, 0]) == int(ids[0, 0])
    assert int(fd["zip"](np.asarray(["no-such-zip"]))[0, 0]) == 0  # unseen value -> the unknown row
    for name in ("timestamp", "age"):
        assert isinstance(fd[name], DenseFeat)
        v = fd[name](df[name].values)
        # min-max scaled; like the reference (inputs.py:308-316) the column is cast to float32 BEFORE scaling, so the
        # 9-digit timestamps land within float32 rounding of [0, 1], not exactly on it
        assert v.dtype == np.float32 and abs(float(v.min())) < 1e-5 and abs(float(v.max()) - 1.0) < 1e-5
    assert isinstance(fd["genres"], MultiValCsvFeat) and fd["genres"].feat_size == len(domains["genres"]) + 1 == 20
    values, offsets = fd["genres"](df.genres.values)
    n_per = np.diff(offsets)
    assert offsets.shape == (len(df) + 1,) and values.min() >= 1 and values.max() <= 19
    assert (n_per == df.genres.str.split("|").str.len().values).all() and 1 <= n_per.min() and n_per.max() <= 6


def _plain_inputs(fd, df):
    """Encoded columns + a feature dictionary without encoders/scalers (what tests/parity_util drives the oracle with)."""
    from recman_b200.th.input import DenseFeat, FeatureDictionary, MultiValCsvFeat, SparseFeat

    plain, X = FeatureDictionary(), {}
    for name, f in fd.items():
        if type(f) is SparseFeat:
            plain[name] = SparseFeat(name, f.feat_size - 1, encoder=False)
            X[name] = f(df[name].values).reshape(-1)
        elif isinstance(f, DenseFeat):
            plain[name] = DenseFeat(name, scaler=False)
            X[name] = f(df[name].values).reshape(-1)
        else:
            plain[name] = MultiValCsvFeat(name, tags=f.tags)
            X[name] = df[name].values
    return plain, X


@pytest.mark.gpu
def test_c1_deepfm_k8_b256_on_ml100k_matches_oracle():
    """Config 1 at its own size: m = 6 embedding fields (5 sparse + genres, sqrtn-pooled), 2 dense, k = 8, B = 256."""
    import torch

    from recman_b200.examples.utils import create_ml_features
    from recman_b200.th import DeepFM
    from recman_b200.th.input import DataInputs
    from tests import parity_util as pu

    df, domains = _sample()
    fd = create_ml_features(df, domains)
    batch = df.iloc[:256]
    plain, X = _plain_inputs(fd, batch)
    # the DataFrame path (encoders, scalers, CSV parsing) produces exactly the tensors the pre-encoded columns do
    a = DataInputs("cuda").load(fd, batch, batch.label.values)
    b = DataInputs("cuda").load(plain, X, batch.label.values)
    assert torch.equal(a.sparse_ids, b.sparse_ids) and torch.equal(a.dense, b.dense)
    assert all(torch.equal(u, v) for u, v in zip(a.csr["genres"], b.csr["genres"]))
    model = DeepFM(plain, embedding_size=8, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=256)
    report = pu.compare(model, X, batch.label.values)
    assert max(report.values()) < 1e-4


@pytest.mark.gpu
def test_c1_fit_on_ml100k_sample_learns():
    """The reference workflow (examples/DeepFM_ml.py): fit on a DataFrame, metrics improve over the untrained model."""
    from recman_b200.examples.utils import create_ml_features
    from recman_b200.th import DeepFM
    from recman_b200.th.metric import LogLoss, RocAucScore

    df, domains = _sample()
    fd = create_ml_features(df, domains)
    train, valid = df.iloc[:2400], df.iloc[2400:]
    model = DeepFM(fd, embedding_size=8, epoch=6, batch_size=256, learning_rate=0.01, deep_dropout=(1, 1, 1),
                   eval_metric=[LogLoss(), RocAucScore()])
    model.fit(train, train.label.values, valid, valid.label.values, random_seed_for_mini_batch=False)
    (ll0, auc0), _ = model.history[0]
    (ll1, auc1), (vll, vauc) = model.history[-1]
    assert ll1 < ll0 - 0.05 and auc1 > 0.75 and auc1 > auc0
    assert np.isfinite(vll) and vauc > 0.5
    pred = model.predict(valid)
    assert pred.shape[0] == len(valid) and float(pred.min()) >= 0.0 and float(pred.max()) <= 1.0


@pytest.mark.gpu
def test_xdeepfm_sweep_with_best_model_finder_on_ml100k_sample(tmp_path):
    """The reference's xDeepFM example workflow (examples/xDeepFM_ml.py): grid search over the learning rate, best model
    kept (and saved) by the epoch callback, restored into a fresh model."""
    import torch

    from recman_b200.examples.xDeepFM_ml import sweep
    from recman_b200.th import xDeepFM

    df, domains = _sample()
    train, valid, test = df.iloc[:2000], df.iloc[2000:2600], df.iloc[2600:]
    finder, fd = sweep(train, valid, test, domains, learning_rates=(0.01, 0.002), epoch=1, batch_size=128,
                       save_model=True, out_dir=str(tmp_path))
    best = finder.best_model
    assert best is not None and best.hparams["learning_rate"] in (0.01, 0.002)
    train_res, valid_res = finder.best_eval_results
    assert finder.best_score == float(valid_res[0]) and np.isfinite(finder.best_score)
    assert (tmp_path / "ckpt_model.pt").exists() and (tmp_path / "hparams").exists() and (tmp_path / "feat_dict").exists()
    pred = best.predict(test)
    assert pred.shape == (len(test),) and np.all((pred > 0) & (pred < 1))
    # restore the saved variables into a fresh model: same predictions
    fresh = xDeepFM(fd, dict(best.hparams), batch_size=128)
    fresh.predict(test.iloc[:4])  # creates the variables
    fresh.restore(str(tmp_path / "ckpt_model.pt"))
    np.testing.assert_allclose(fresh.predict(test), pred, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_dcn_example_on_ml100k_sample():
    """recman/examples/DCN_ml.py on the sample: DCN with the multi-valued genres field (sqrtn-pooled lookup) trains."""
    from recman_b200.examples.DCN_ml import train

    df, domains = _sample()
    model, _ = train(df.iloc[:2400], df.iloc[2400:2700], df.iloc[2700:], domains, epoch=3, batch_size=128,
                     learning_rate=0.01)
    (ll0, auc0), _ = model.history[0]
    (ll1, auc1), (vll, vauc) = model.history[-1]
    assert ll1 < ll0 and auc1 > auc0 and np.isfinite(vll) and 0.0 <= vauc <= 1.0
    pred = model.predict(df.iloc[2700:])
    assert pred.shape == (300,) and np.all((pred >= 0) & (pred <= 1))
