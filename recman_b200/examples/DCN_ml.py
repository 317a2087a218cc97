"""DCN on MovieLens-100k - the workflow of recman/examples/DCN_ml.py (3 cross layers + (32, 32) MLP by default).

    python -m recman_b200.examples.DCN_ml --data-dir /path/to/data [--frac 0.2] [--epoch 5]
"""
import argparse

import pandas as pd


def train(df_train, df_valid, df_test, domains, epoch=5, batch_size=64, learning_rate=0.001):
    """-> (model, feat_dict); importable so that the tests can run it on the bundled sample."""
    from ..th import DCN
    from ..th.metric import LogLoss, RocAucScore
    from .utils import create_ml_features

    feat_dict = create_ml_features(pd.concat([df_train, df_valid, df_test], axis=0), domains)
    model = DCN(feat_dict, learning_rate=learning_rate, epoch=epoch, batch_size=batch_size,
                eval_metric=[LogLoss(), RocAucScore()])
    model.fit(df_train, df_train["label"].values, df_valid, df_valid["label"].values)
    return model, feat_dict


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--frac", type=float, default=0.2)
    ap.add_argument("--epoch", type=int, default=5)
    args = ap.parse_args(argv)
    from .utils import get_ml_dataset

    df_train, df_valid, df_test, domains = get_ml_dataset(args.data_dir, frac=args.frac)
    model, _ = train(df_train, df_valid, df_test, domains, epoch=args.epoch)
    res = model.evaluate(df_test, df_test["label"].values)
    print("test", [(str(f), round(float(r), 4)) for f, r in zip(model.metrics, res)])
    return model


if __name__ == "__main__":
    main()
