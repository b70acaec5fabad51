"""Dense U V^T top-n recommendation scoring (evaluation): the path's only tensor-core user (``pmf_topn``).

No reference code exists for this (BASELINE north_star item 3); semantics are fixed by
``oracle/pmf_oracle.py::topn``: float32 scores accumulated in k order, ranking (score desc, item index asc).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi
from ._engine import pad_table, row_stride


def _as_table(F, device):
    """(R, K) NumPy / torch -> (float32 CUDA tensor [R, ld], K)."""
    if isinstance(F, torch.Tensor):
        K = F.shape[1]
        ld = row_stride(K)
        if F.is_cuda and F.dtype == torch.float32 and F.is_contiguous() and K == ld:
            return F, K
        out = torch.zeros((F.shape[0], ld), dtype=torch.float32, device=device)
        out[:, :K] = F.to(device=device, dtype=torch.float32)
        return out, K
    F = np.asarray(F)
    return pad_table(F, row_stride(F.shape[1]), device), F.shape[1]


def top_n(F_user, F_item, n=50, user_rows=None, tensor_cores=True, batch_rows=None, device=None, return_stats=False):
    """Top-``n`` items for every row of ``F_user`` (or the rows listed in ``user_rows``).

    Returns ``(idx int32[B, n], score float32[B, n])`` as NumPy arrays.  ``tensor_cores=True`` scores with
    tcgen05 (bf16 operands) and re-scores a provably sufficient candidate set exactly, so the indices equal
    the exact path's bit for bit; for n <= 256 and K <= 160 the scores are filtered in the MMA epilogue and
    never written to HBM.  ``tensor_cores="unfused"`` keeps the score matrix in HBM (comparison point),
    ``False`` scores exactly on CUDA cores.  ``batch_rows``: user rows per library call (default 8192 fused,
    1024 otherwise -- the unfused modes need 4*n_items bytes of workspace per row).
    """
    mode = 2 if tensor_cores == "unfused" else int(bool(tensor_cores))
    _cabi.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    Fu, K = _as_table(F_user, device)
    Fi, Ki = _as_table(F_item, device)
    if K != Ki:
        raise ValueError("factor tables disagree on K")
    ld = Fu.shape[1]
    rows = None
    B = Fu.shape[0]
    if user_rows is not None:
        rows = torch.as_tensor(np.asarray(user_rows, dtype=np.int32)).to(device)
        B = rows.numel()
    M = Fi.shape[0]
    idx = torch.empty((B, n), dtype=torch.int32, device=device)
    score = torch.empty((B, n), dtype=torch.float32, device=device)
    stats_total = np.zeros(2, dtype=np.int64)
    stats = torch.zeros(2, dtype=torch.int32, device=device)
    lib = _cabi.load()
    if batch_rows is None:
        batch_rows = 8192 if lib.pmf_topn_workspace_bytes_ex(256, M, K, n, mode) < 256 * M * 4 else 1024
    chunk = max(128, min(int(batch_rows), max(B, 1)))
    ws_bytes = lib.pmf_topn_workspace_bytes_ex(chunk, M, K, n, mode)
    if ws_bytes < 0:
        raise ValueError("bad top-n shape")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        for s in range(0, B, chunk):
            e = min(s + chunk, B)
            if rows is not None:
                base_ptr, rows_ptr = Fu.data_ptr(), rows[s:e].data_ptr()
            else:
                base_ptr, rows_ptr = Fu[s:e].data_ptr(), None
            _cabi.call("pmf_topn", base_ptr, rows_ptr, e - s, Fi.data_ptr(), M, K, ld, n, mode,
                       idx[s:e].data_ptr(), score[s:e].data_ptr(), ws.data_ptr(), ws_bytes, stats.data_ptr(),
                       _cabi.stream_ptr())
            if return_stats:
                stats_total += stats.cpu().numpy()
    out = (idx.cpu().numpy(), score.cpu().numpy())
    return out + ({"exact_fallback_rows": int(stats_total[0]), "candidates_rescored": int(stats_total[1])},) if return_stats else out


def recommend(model, user_ids, n=50, tensor_cores=True):
    """Top-``n`` items for ``user_ids`` from a fitted model's factors (E_theta/E_beta, m_theta/m_beta or softplus params)."""
    name = type(model).__name__
    if name == "HPF_PyTorch":
        Fu, Fi = model.theta.detach(), model.beta.detach()
    elif name == "GaussianMFCAVI":
        e = model._engine
        Fu, Fi = e.m_theta, e.m_beta
    else:
        e = model._engine
        Fu, Fi = e.E_theta, e.E_beta
    if Fu.shape[1] != row_stride(Fu.shape[1]) or name == "HPF_PyTorch":
        return top_n(Fu, Fi, n, user_rows=user_ids, tensor_cores=tensor_cores)
    K = model.config.n_factors
    return top_n(Fu[:, :K] if Fu.shape[1] != K else Fu, Fi[:, :K] if Fi.shape[1] != K else Fi, n, user_rows=user_ids,
                 tensor_cores=tensor_cores)
