"""Make tests/golden/ml100k_sample.csv.gz: a seeded 3000-row sample of the merged ua.base frame of the MovieLens-100k
copy the reference bundles (data/ml-100k), produced by this repo's loader, plus the genre / occupation domains.

Run in the development container only (the GPU box has no /root/reference):
    python tests/golden/make_ml100k_sample.py [/root/reference/data]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import importlib.util

    # load the loader module by path: importing the package would pull in the CUDA library, which is not needed here
    spec = importlib.util.spec_from_file_location(
        "ml_100k", os.path.join(ROOT, "recman_b200", "examples", "datasets", "ml_100k.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    data_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    df_train, df_test, domains = mod.get_data(data_dir)
    sample = df_train.sample(n=3000, random_state=2019).reset_index(drop=True)
    out = os.path.join(ROOT, "tests", "golden")
    sample.drop(columns=["title"]).to_csv(os.path.join(out, "ml100k_sample.csv.gz"), index=False, compression="gzip")
    meta = dict(domains=domains, n_train_all=int(len(df_train)), n_test_all=int(len(df_test)),
                n_users=int(df_train.user_id.nunique()), n_items=int(df_train.item_id.nunique()),
                genres_mean_len=float(df_train.genres.str.split("|").str.len().mean()),
                columns=list(sample.columns))
    json.dump(meta, open(os.path.join(out, "ml100k_sample_meta.json"), "w"), indent=1)
    print(len(sample), "rows;", meta["n_train_all"], "train rows in ua.base;", meta["n_users"], "users;",
          meta["n_items"], "items; mean genres/item-interaction", round(meta["genres_mean_len"], 3))


if __name__ == "__main__":
    main()
