"""BASELINE config 1: DeepFM (k=8, batch 256) on MovieLens-100k, the workflow of recman/examples/DeepFM_ml.py.

    python -m recman_b200.examples.DeepFM_ml --data-dir /path/to/data [--frac 0.2] [--epoch 5]

`--data-dir` is the directory that holds `ml-100k/` (the reference bundles it under data/).
"""
import argparse

import pandas as pd


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--frac", type=float, default=0.2)
    ap.add_argument("--epoch", type=int, default=5)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--embedding-size", type=int, default=8)
    args = ap.parse_args(argv)

    from ..th import DeepFM
    from ..th.metric import LogLoss, RocAucScore
    from .utils import create_ml_features, get_ml_dataset

    df_train, df_valid, df_test, domains = get_ml_dataset(args.data_dir, frac=args.frac)
    feat_dict = create_ml_features(pd.concat([df_train, df_valid, df_test], axis=0), domains)
    model = DeepFM(feat_dict, embedding_size=args.embedding_size, learning_rate=0.001, epoch=args.epoch,
                   batch_size=args.batch_size, eval_metric=[LogLoss(), RocAucScore()])
    model.fit(df_train, df_train["label"].values, df_valid, df_valid["label"].values)
    for name, df in (("valid", df_valid), ("test", df_test)):
        res = model.evaluate(df, df["label"].values)
        print(name, [(str(f), round(float(r), 4)) for f, r in zip(model.metrics, res)])
    return model


if __name__ == "__main__":
    main()
