"""Dataset split + feature dictionary of the ml-100k examples (recman/examples/utils.py:9-75)."""
import numpy as np

from ..th.input import DenseFeat, FeatureDictionary, MultiValCsvFeat, SparseFeat
from .datasets.ml_100k import get_data


def add_labels(df):
    """rating >= 4 -> 1, else 0 (utils.py:14-17)."""
    df = df.copy()
    df["label"] = (df.rating >= 4).astype(np.float32)
    return df


def get_ml_dataset(data_dir, frac=0.5, random_seed=2019):
    """-> (df_train, df_valid, df_test, domains): `frac` of ua.base, split 70/30, plus ua.test (utils.py:9-27)."""
    df_all, df_test, domains = get_data(data_dir)
    df_all = add_labels(df_all.sample(frac=frac, random_state=random_seed))
    df_test = add_labels(df_test)
    df_train = df_all.sample(frac=0.7, random_state=random_seed)
    df_valid = df_all.drop(df_train.index)
    return df_train, df_valid, df_test, domains


def create_ml_features(df_data, domains):
    """user_id, item_id, gender, occupation, zip (sparse), timestamp, age (dense, min-max), genres (multi-valued);
    the same eight features in the same order as utils.py:30-75."""
    from sklearn.preprocessing import MinMaxScaler

    fd = FeatureDictionary()
    for name in ("user_id", "item_id", "gender", "occupation", "zip"):
        fd[name] = SparseFeat(name=name, feat_size=len(np.unique(df_data[name].values)))
    fd["timestamp"] = DenseFeat(name="timestamp", scaler=MinMaxScaler())
    fd["age"] = DenseFeat(name="age", scaler=MinMaxScaler())
    fd["genres"] = MultiValCsvFeat(name="genres", tags=domains["genres"])
    fd.initialize(df_data)
    return fd
