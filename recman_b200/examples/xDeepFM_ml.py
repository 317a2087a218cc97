"""xDeepFM on MovieLens-100k with a small hyper-parameter sweep - the workflow of recman/examples/xDeepFM_ml.py
(grid search over learning_rate / optimizer, best model kept by BestModelFinder; TensorBoard logging is out of scope).

    python -m recman_b200.examples.xDeepFM_ml --data-dir /path/to/data [--frac 0.7] [--epoch 1]
"""
import argparse
import logging

import pandas as pd

log = logging.getLogger(__name__)
RANDOM_SEED = 2019


def sweep(df_train, df_valid, df_test, domains, learning_rates=(0.01, 0.005), optimizers=("adam",), epoch=1,
          batch_size=128, save_model=False, out_dir="."):
    """-> (BestModelFinder, feat_dict).  Importable so that the tests can run it on the bundled sample."""
    from ..th import xDeepFM
    from ..th.BestModelFinder import BestModelFinder
    from ..th.hparams import Discrete
    from ..th.hparams import xDeepFM as HyperParams
    from ..th.metric import LogLoss, RocAucScore
    from .utils import create_ml_features

    df_all = pd.concat([df_train, df_valid, df_test], axis=0)
    feat_dict = create_ml_features(df_all, domains)
    hp_params = HyperParams()
    hp_params["learning_rate"](Discrete(list(learning_rates)))
    hp_params["optimizer"](Discrete(list(optimizers)))
    finder = BestModelFinder(save_model=save_model, out_dir=out_dir)
    for hp_val in hp_params.grid_search():
        model = xDeepFM(feat_dict, hp_val, batch_size=batch_size, metrics=(LogLoss(), RocAucScore()),
                        random_seed=RANDOM_SEED, epoch=epoch)
        model.fit(X_train=df_train, y_train=df_train["label"].values, X_valid=df_valid,
                  y_valid=df_valid["label"].values, epoch_callback=finder, random_seed_for_mini_batch=False)
    return finder, feat_dict


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--frac", type=float, default=0.7)
    ap.add_argument("--epoch", type=int, default=1)
    args = ap.parse_args(argv)
    from .utils import get_ml_dataset

    df_train, df_valid, df_test, domains = get_ml_dataset(args.data_dir, frac=args.frac)
    finder, _ = sweep(df_train, df_valid, df_test, domains, epoch=args.epoch)
    best = finder.best_model
    log.info("best validation log-loss %.4f with lr=%s", finder.best_score, best.hparams["learning_rate"])
    res = best.evaluate(df_test, df_test["label"].values)
    print("test", [(str(f), round(float(r), 4)) for f, r in zip(best.metrics, res)])
    return finder


if __name__ == "__main__":
    main()
