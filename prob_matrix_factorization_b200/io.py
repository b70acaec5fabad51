"""``data/embeddings/<model>/`` and ``data/predictions/<model>/`` writers with the reference's CSV layout.

The reference's ``train_*_full.py`` scripts write these files themselves from the model's public arrays
(train_poisson_full.py:62-123, train_gaussian_full.py:71-137, train_hpf_cavi_full.py:71-139,
train_hpf_pytorch_full.py:112-171); because the drop-in classes expose the same arrays those scripts work
unchanged.  This module is the same writer for callers that do not go through the scripts; the
``src/analysis`` readers (analyze_top_dimensions.py:19-43, embedding_viz.py:30-35) consume its output.
"""
from __future__ import annotations

import os
from dataclasses import asdict

import numpy as np
import pandas as pd

MODEL_DIRS = {"PoissonMFCAVI": "poisson_mf", "HPF_CAVI": "hpf_cavi", "GaussianMFCAVI": "gaussian_mf",
              "HPF_PyTorch": "hpf_pytorch",
              "PoissonMFExtendedCAVI": "poisson_mf_extended"}   # no reference script writes this one; same layout


def embeddings_of(model):
    """(user_emb, item_emb) exactly as each reference script picks them."""
    name = type(model).__name__
    if name == "GaussianMFCAVI":
        return model.m_theta, model.m_beta                                  # train_gaussian_full.py:77-80
    if name == "HPF_PyTorch":
        return model.theta.detach().cpu().numpy(), model.beta.detach().cpu().numpy()   # train_hpf_pytorch_full.py:119-120
    return model.E_theta, model.E_beta                                      # train_poisson_full.py:68-76


def save_embeddings(model, out_root="data", recipe_ids=None, global_mean=None):
    sub = MODEL_DIRS[type(model).__name__]
    out_dir = os.path.join(out_root, "embeddings", sub)
    os.makedirs(out_dir, exist_ok=True)
    user_emb, item_emb = embeddings_of(model)
    pd.DataFrame(user_emb).to_csv(os.path.join(out_dir, "user_embeddings.csv"), index=False)
    item_df = pd.DataFrame(item_emb)
    if recipe_ids is not None and len(recipe_ids) == len(item_df):
        item_df.insert(0, "recipe_id", recipe_ids)                          # train_poisson_full.py:83-91
    item_df.to_csv(os.path.join(out_dir, "item_embeddings.csv"), index=False)
    with open(os.path.join(out_dir, "config.txt"), "w") as f:
        f.write(str(asdict(model.config)))
        if global_mean is not None:
            f.write(f"\nglobal_mean: {global_mean}")                        # train_gaussian_full.py:106
    return out_dir


def save_test_predictions(model, test_u, test_i, y_true, y_pred, out_root="data"):
    sub = MODEL_DIRS[type(model).__name__]
    pred_dir = os.path.join(out_root, "predictions", sub)
    os.makedirs(pred_dir, exist_ok=True)
    pd.DataFrame({"u": np.asarray(test_u), "i": np.asarray(test_i), "y_true": np.asarray(y_true),
                  "y_pred": np.asarray(y_pred)}).to_csv(os.path.join(pred_dir, "test_predictions.csv"), index=False)
    return pred_dir
