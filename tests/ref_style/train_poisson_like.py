"""A driver in the style of the reference's train scripts (train_poisson_full.py:15-123; TEST INFRASTRUCTURE, not a copy):
it only knows the reference's import paths and call surface --

    from src.models.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    from src.evaluation.metrics import rmse, macro_mae
    import matplotlib.pyplot as plt                      (compare_models.py:4 imports it at module level)

-- reads data/processed/interactions_{train,validation,test}.csv (load_data.py:93-113), takes its hyper-parameters from a
best_hyperparams.txt dict-literal line (compare_models.py:25-47), fits, and writes data/embeddings/poisson_mf/* and
data/predictions/poisson_mf/test_predictions.csv exactly as the reference scripts do.  Run on top of
prob_matrix_factorization_b200.dropin it exercises the drop-in boundary end to end.
"""
import ast
import os
from dataclasses import asdict

import matplotlib.pyplot as plt  # noqa: F401  (stubbed where matplotlib is not installed)
import pandas as pd

from src.evaluation.metrics import macro_mae, rmse
from src.models.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig


def load_interactions(path):
    df = pd.read_csv(path)
    return df[["u", "i", "rating"]]


def load_best_hyperparams(path="best_hyperparams.txt"):
    out = {}
    with open(path) as f:
        for line in f:
            if ":" in line and "{" in line:
                name, literal = line.split(":", 1)
                out[name.strip()] = ast.literal_eval(literal.strip())
    return out


def main(dataset_mode="train+val"):
    train_df, val_df, test_df = (load_interactions(f"data/processed/interactions_{s}.csv") for s in ("train", "validation", "test"))
    df = pd.concat([train_df, val_df])[["u", "i", "rating"]] if dataset_mode == "train+val" else train_df
    config = PoissonMFCAVIConfig(**load_best_hyperparams()["PoissonMF"])
    model = PoissonMFCAVI(config)
    model.fit(df)
    out_dir = "data/embeddings/poisson_mf"
    os.makedirs(out_dir, exist_ok=True)
    pd.DataFrame(model.E_theta).to_csv(os.path.join(out_dir, "user_embeddings.csv"), index=False)
    pd.DataFrame(model.E_beta).to_csv(os.path.join(out_dir, "item_embeddings.csv"), index=False)
    with open(os.path.join(out_dir, "config.txt"), "w") as f:
        f.write(str(asdict(config)))
    pred_dir = "data/predictions/poisson_mf"
    os.makedirs(pred_dir, exist_ok=True)
    test_u, test_i, y_true = test_df["u"].to_numpy(), test_df["i"].to_numpy(), test_df["rating"].to_numpy()
    y_pred = model.predict(test_u, test_i)
    print(f"Test Set Metrics: MacroMAE={macro_mae(y_true, y_pred):.4f} | RMSE={rmse(y_true, y_pred):.4f}")
    pd.DataFrame({"u": test_u, "i": test_i, "y_true": y_true, "y_pred": y_pred}).to_csv(
        os.path.join(pred_dir, "test_predictions.csv"), index=False)
    plt.figure()
    plt.plot([0, 1], [0, 1])
    plt.savefig("unused.png")


if __name__ == "__main__":
    main()
