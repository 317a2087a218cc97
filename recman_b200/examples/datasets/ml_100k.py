"""MovieLens-100k loader: the merged (interaction, user, item) frames of recman/examples/datasets/ml_100k.py:4-96.

Same file set, column names, merge keys and return value ``(df_train_all, df_test_all, domains)`` as the reference's
``get_data``; the per-row ``DataFrame.apply`` that builds the ``"a|b|d"`` genre strings is replaced by one vectorised
pass over the 19 indicator columns.
"""
import os

import numpy as np
import pandas as pd

INTERACTION_COLS = ["user_id", "item_id", "rating", "timestamp"]
USER_COLS = ["user_id", "age", "gender", "occupation", "zip"]
ITEM_HEAD_COLS = ["item_id", "title", "release_date", "video_release_date", "imdb_url"]


def _read(path, sep, names):
    return pd.read_csv(path, delimiter=sep, header=None, encoding="latin-1", names=names)


def genre_strings(df_items: pd.DataFrame, genres) -> np.ndarray:
    """One ``"Action|Comedy"`` string per item from the 0/1 indicator columns, genres in u.genre order."""
    flags = df_items[list(genres)].to_numpy() == 1
    names = np.asarray(list(genres), dtype=object)
    return np.asarray(["|".join(names[row]) for row in flags], dtype=object)


def get_data(data_dir, file_set="a"):
    root = os.path.join(data_dir, "ml-100k")
    df_genres = _read(os.path.join(root, "u.genre"), "|", ["genre", "id"])
    df_occupations = _read(os.path.join(root, "u.occupation"), "|", ["occupation"])
    df_users = _read(os.path.join(root, "u.user"), "|", USER_COLS)
    genres = df_genres.genre.unique().tolist()
    df_items = _read(os.path.join(root, "u.item"), "|", ITEM_HEAD_COLS + genres)
    df_items["genres"] = genre_strings(df_items, genres)
    item_cols = df_items[["item_id", "title", "release_date", "genres"]]

    def merged(split):
        inter = _read(os.path.join(root, f"u{file_set}.{split}"), "\t", INTERACTION_COLS)
        return inter.merge(df_users, on="user_id").merge(item_cols, on="item_id")

    domains = dict(genres=df_genres.genre.tolist(), occupations=df_occupations.occupation.tolist())
    return merged("base"), merged("test"), domains
