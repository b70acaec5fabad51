"""Drop-in for the reference's ``src.models.hpf_pytorch`` (hpf_pytorch.py:9-195): gradient-based (MAP)
hierarchical Poisson factorisation as a ``torch.nn.Module``.

The scripts that drive it (compare_models.py:287-313, train_hpf_pytorch_full.py:76-108,
tune_all_models.py:238-266) build ``torch.optim.Adam(model.parameters())``, feed CPU ``LongTensor`` /
``FloatTensor`` mini-batches and call ``loss(...).backward()``; they never call ``.to(device)``.  This
class therefore owns CUDA parameters internally, accepts CPU batches, and implements ``loss`` as a
``torch.autograd.Function`` whose forward launches ONE fused kernel (``pmf_hpf_map_loss_grad``) that
computes the loss and the analytic gradient of every term on the touched rows only; ``backward`` hands
those gradients to autograd, so the scripts' own optimiser loop runs unchanged.

``fit_epochs`` is the loader-free fast path: it replays torch's DataLoader shuffle stream
(bit-identical permutations) and steps with the fused ``pmf_adam_dense_step`` kernel.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi


LAZY_LAUNCHES_PER_STEP = 1     # pmf_hpf_map_lazy_epoch: one fused settle + loss + gradient kernel per mini-batch

@dataclass
class HPF_PyTorch_Config:
    n_factors: int = 20
    a: float = 0.3
    a_prime: float = 1.0
    b_prime: float = 1.0
    c: float = 0.3
    c_prime: float = 1.0
    d_prime: float = 1.0
    lr: float = 0.001
    batch_size: int = 1024
    epochs: int = 20
    device: str = "cpu"        # kept for config round-trips; the engine always runs on the GPU
    verbose: bool = True


def _loss_grad(mod, users, items, ratings, grads, loss_acc):
    cfg = mod.config
    id_bytes = 8 if users.dtype == torch.int64 else 4
    with torch.cuda.device(mod.theta_uncons.device):
        _cabi.call("pmf_hpf_map_loss_grad", users.data_ptr(), items.data_ptr(), id_bytes, ratings.data_ptr(),
                   users.numel(), mod.theta_uncons.data_ptr(), mod.beta_uncons.data_ptr(), mod.xi_uncons.data_ptr(),
                   mod.eta_uncons.data_ptr(), mod.user_scale.data_ptr(), mod.item_scale.data_ptr(), mod.n_users,
                   mod.n_items, mod.K, cfg.a, cfg.a_prime, cfg.b_prime, cfg.c, cfg.c_prime, cfg.d_prime,
                   grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), grads[3].data_ptr(),
                   loss_acc.data_ptr(), mod._bad.data_ptr(), _cabi.stream_ptr())


class _FusedLoss(torch.autograd.Function):
    """loss = HPF_PyTorch.loss(batch); gradients come from the same kernel launch as the value."""

    @staticmethod
    def forward(ctx, theta, beta, xi, eta, mod, users, items, ratings):
        grads = [torch.zeros_like(p) for p in (theta, beta, xi, eta)]
        acc = torch.zeros((), dtype=torch.float64, device=theta.device)
        _loss_grad(mod, users, items, ratings, grads, acc)
        ctx.grads = grads
        return acc.to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        g = ctx.grads
        ctx.grads = None
        return g[0].mul_(gout), g[1].mul_(gout), g[2].mul_(gout), g[3].mul_(gout), None, None, None, None


class _EpochPermutations:
    """The DataLoader's shuffles, replayed bit for bit, off the critical path.

    Per epoch the reference's loop draws ``_base_seed`` (dataloader.py ``_BaseDataLoaderIter.__init__``) and the sampler's
    ``seed`` (sampler.py ``RandomSampler.__iter__``) from the global CPU generator and shuffles with
    ``torch.randperm(n, generator=Generator().manual_seed(seed))``.  The two draws are cheap and sequential; the
    permutation itself is a serial Fisher-Yates on the host (20-100 ms for 1.1 M ratings -- several times the GPU time
    of the epoch it feeds), but permutations of different epochs only depend on their seeds, so they are computed by a
    few host threads ahead of the GPU.  ``prefetch=False`` draws and shuffles epoch by epoch (needed when a callback
    may itself use the global generator between epochs)."""

    def __init__(self, n, epochs, device, prefetch=True, workers=8):
        self.n, self.epochs, self.device = n, epochs, device
        self.prefetch = prefetch and epochs > 1
        self._futures = []
        if self.prefetch:
            from concurrent.futures import ThreadPoolExecutor
            seeds = [self._draw_seed() for _ in range(epochs)]
            self._pool = ThreadPoolExecutor(max_workers=min(workers, epochs))
            self._futures = [self._pool.submit(self._permute, s) for s in seeds]

    @staticmethod
    def _draw_seed():
        torch.empty((), dtype=torch.int64).random_()                         # _base_seed (value unused, draw kept)
        return int(torch.empty((), dtype=torch.int64).random_().item())      # RandomSampler's seed

    def _permute(self, seed):
        gen = torch.Generator()
        gen.manual_seed(seed)
        perm = torch.randperm(self.n, generator=gen)
        return perm.pin_memory() if torch.device(self.device).type == "cuda" else perm

    def __iter__(self):
        for ep in range(self.epochs):
            perm = self._futures[ep].result() if self.prefetch else self._permute(self._draw_seed())
            if self.prefetch:
                self._futures[ep] = None
            yield perm.to(self.device, non_blocking=True)
        if self.prefetch:
            self._pool.shutdown(wait=False)


class HPF_PyTorch(nn.Module):
    def __init__(self, n_users, n_items, user_counts, item_counts, config: HPF_PyTorch_Config, device=None):
        super().__init__()
        _cabi.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.config = config
        self.n_users = int(n_users)
        self.n_items = int(n_items)
        self.K = int(config.n_factors)
        # Buffers for scaling (hpf_pytorch.py:34-35), computed on the host exactly as the reference does
        self.register_buffer("user_scale", (1.0 / (torch.tensor(user_counts, dtype=torch.float32) + 1e-6)).to(dev))
        self.register_buffer("item_scale", (1.0 / (torch.tensor(item_counts, dtype=torch.float32) + 1e-6)).to(dev))
        # Same draws from the global CPU generator, same order (theta, beta, xi, eta; hpf_pytorch.py:39-48)
        self.theta_uncons = nn.Parameter((torch.randn(self.n_users, self.K) * 0.1).to(dev))
        self.beta_uncons = nn.Parameter((torch.randn(self.n_items, self.K) * 0.1).to(dev))
        self.xi_uncons = nn.Parameter((torch.randn(self.n_users) * 0.1).to(dev))
        self.eta_uncons = nn.Parameter((torch.randn(self.n_items) * 0.1).to(dev))
        self.register_buffer("_bad", torch.zeros(1, dtype=torch.int32, device=dev), persistent=False)
        self._adam = None

    # -- constrained views (outputs only; the training kernels apply softplus on touched rows) -------
    @property
    def theta(self):
        return F.softplus(self.theta_uncons)

    @property
    def beta(self):
        return F.softplus(self.beta_uncons)

    @property
    def xi(self):
        return F.softplus(self.xi_uncons)

    @property
    def eta(self):
        return F.softplus(self.eta_uncons)

    def _ids(self, t):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        if t.dtype not in (torch.int64, torch.int32):
            t = t.long()
        return t.to(self.theta_uncons.device, non_blocking=True).contiguous()

    def forward(self, user_ids, item_ids):
        u, i = self._ids(user_ids), self._ids(item_ids)
        out = torch.empty(u.numel(), dtype=torch.float32, device=u.device)
        with torch.cuda.device(u.device):
            _cabi.call("pmf_hpf_map_predict", u.data_ptr(), i.data_ptr(), 8 if u.dtype == torch.int64 else 4, u.numel(),
                       self.theta_uncons.data_ptr(), self.beta_uncons.data_ptr(), self.n_users, self.n_items, self.K,
                       out.data_ptr(), _cabi.stream_ptr())
        return out

    def loss(self, user_ids, item_ids, ratings):
        """Negative log joint of the mini-batch (sum over the batch), differentiable w.r.t. the parameters."""
        u, i = self._ids(user_ids), self._ids(item_ids)
        if u.dtype != i.dtype:
            u, i = u.long(), i.long()
        r = torch.as_tensor(ratings).to(torch.float32).to(u.device, non_blocking=True).contiguous()
        return _FusedLoss.apply(self.theta_uncons, self.beta_uncons, self.xi_uncons, self.eta_uncons, self, u, i, r)

    def predict(self, user_ids, item_ids):
        with torch.no_grad():
            preds = self.forward(user_ids, item_ids)
        return preds.cpu().numpy()

    def check_ids(self):
        """Raise like torch indexing would if any batch so far carried an out-of-range id (one D2H)."""
        if int(self._bad.item()) != 0:
            raise IndexError("index out of range in HPF_PyTorch batch")

    # -- loader-free training ------------------------------------------------------------------------
    def fit_epochs(self, users, items, ratings, epochs=None, batch_size=4096, lr=None, shuffle=True, on_epoch=None,
                   lazy=False, stats=None):
        """The scripts' loop (compare_models.py:299-313) without the DataLoader.

        Per epoch the shuffle is torch's own: ``DataLoader.__iter__`` draws ``_base_seed`` then
        ``RandomSampler`` draws its seed from the global CPU generator and calls
        ``torch.randperm(n, generator=Generator().manual_seed(seed))`` -- replayed here bit for bit, so a
        run under the same ``torch.manual_seed`` visits the same mini-batches as the reference loop.
        ``lazy=False`` (default): one fused loss+gradient kernel and one dense Adam kernel per tensor per step.
        ``lazy=True``: touch-only Adam (``pmf_hpf_map_lazy_epoch``): ONE kernel per mini-batch settles the rows the batch
        references (their deferred Adam step + the zero-gradient steps they skipped, in closed form), then adds the
        batch's loss and gradients; ~250x fewer bytes than the dense update and the same result up to float32 rounding
        (tests: 1e-5 max-norm relative on parameters and moments).  Returns the list of epoch
        losses (sum of mini-batch losses, as the scripts print).  ``stats`` (a dict) receives ``device_ms`` -- CUDA-event
        time of the epochs, uploads excluded -- and ``launches`` (this library's kernels).
        """
        cfg = self.config
        epochs = cfg.epochs if epochs is None else epochs
        lr = cfg.lr if lr is None else lr
        dev = self.theta_uncons.device
        u_all, i_all = self._ids(users), self._ids(items)
        r_all = torch.as_tensor(np.asarray(ratings, dtype=np.float32)).to(dev)
        n = u_all.numel()
        params = [self.theta_uncons, self.beta_uncons, self.xi_uncons, self.eta_uncons]
        if self._adam is None:
            self._adam = {"step": 0, "m": [torch.zeros_like(p) for p in params], "v": [torch.zeros_like(p) for p in params]}
        st = self._adam
        grads = st.setdefault("g", [torch.zeros_like(p) for p in params])
        beta1, beta2, eps = 0.9, 0.999, 1e-8
        losses = []
        # device time of the epochs, from the moment the first epoch's permutation is on its way to the GPU
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._ev0 = ev0
        steps_per_epoch = (n + batch_size - 1) // batch_size
        if lazy:
            losses = self._fit_epochs_lazy(u_all, i_all, r_all, epochs, batch_size, lr, shuffle, on_epoch, st, params,
                                           grads, (beta1, beta2, eps))
            ev1.record(torch.cuda.current_stream(dev))
            if stats is not None:
                torch.cuda.synchronize(dev)
                stats.update(device_ms=ev0.elapsed_time(ev1), launches=epochs * (steps_per_epoch * LAZY_LAUNCHES_PER_STEP + 1))
            return losses
        with torch.cuda.device(dev), torch.no_grad():
            perms = iter(_EpochPermutations(n, epochs, dev, prefetch=on_epoch is None)) if shuffle else None
            for ep in range(epochs):
                if shuffle:
                    perm = next(perms)
                    if ep == 0:
                        ev0.record(torch.cuda.current_stream(dev))
                    u_ep, i_ep, r_ep = u_all[perm], i_all[perm], r_all[perm]
                else:
                    if ep == 0:
                        ev0.record(torch.cuda.current_stream(dev))
                    u_ep, i_ep, r_ep = u_all, i_all, r_all
                acc = torch.zeros((), dtype=torch.float64, device=dev)
                for s in range(0, n, batch_size):
                    e = min(s + batch_size, n)
                    for g in grads:
                        g.zero_()
                    _loss_grad(self, u_ep[s:e], i_ep[s:e], r_ep[s:e], grads, acc)
                    st["step"] += 1
                    t = st["step"]
                    step_size = lr / (1.0 - beta1 ** t)
                    bc2_sqrt = math.sqrt(1.0 - beta2 ** t)
                    for p, g, m, v in zip(params, grads, st["m"], st["v"]):
                        _cabi.call("pmf_adam_dense_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                   beta1, beta2, eps, step_size, bc2_sqrt, _cabi.stream_ptr())
                losses.append(float(acc.item()))
                if on_epoch is not None:
                    on_epoch(ep, losses[-1])
        ev1.record(torch.cuda.current_stream(dev))
        if stats is not None:
            torch.cuda.synchronize(dev)
            stats.update(device_ms=ev0.elapsed_time(ev1), launches=epochs * steps_per_epoch * 5)   # loss+grad, 4 x Adam
        self.check_ids()
        return losses

    def _fit_epochs_lazy(self, u_all, i_all, r_all, epochs, batch_size, lr, shuffle, on_epoch, st, params, grads, hyp,
                         closed_form=True):
        import ctypes as C
        beta1, beta2, eps = hyp
        cfg, dev, n = self.config, self.theta_uncons.device, u_all.numel()
        steps_per_epoch = (n + batch_size - 1) // batch_size
        total = st["step"] + epochs * steps_per_epoch
        # per-step scalars exactly as torch computes them (float64 on the host, then float32)
        s_idx = np.arange(total + 1, dtype=np.float64)
        with np.errstate(divide="ignore"):
            step_size = (lr / (1.0 - beta1 ** s_idx)).astype(np.float32)
        bc2 = np.sqrt(1.0 - beta2 ** s_idx).astype(np.float32)
        step_size[0], bc2[0] = 0.0, 1.0
        tab = (torch.from_numpy(step_size).to(dev), torch.from_numpy(bc2).to(dev))
        # closed-form catch-up tables (include/pmf_b200.h pmf_lazy_adam): backward recurrences tail[s] = r (c[s+1] + tail[s+1])
        tails = []
        if closed_form:
            for c, r in ((step_size.astype(np.float64) * bc2, beta1 / math.sqrt(beta2)),
                         (step_size.astype(np.float64) * bc2.astype(np.float64) ** 2, beta1 / beta2)):
                tail = np.zeros(total + 2, dtype=np.float64)
                for s_ in range(total - 1, -1, -1):
                    tail[s_] = r * (c[s_ + 1] + tail[s_ + 1])
                tails.append(torch.from_numpy(tail).to(dev))
            Jr = np.arange(total + 2, dtype=np.float64)
            pow5 = np.stack([beta1 ** Jr, beta2 ** Jr, (beta1 / math.sqrt(beta2)) ** Jr, (beta1 / beta2) ** Jr, beta2 ** (-0.5 * Jr)])
            tails.append(torch.from_numpy(np.ascontiguousarray(pow5)).to(dev))
        i32 = lambda k: torch.zeros(k, dtype=torch.int32, device=dev)
        if "last_user" not in st:
            st.update(last_user=i32(self.n_users), last_item=i32(self.n_items), claim_user=i32(self.n_users),
                      claim_item=i32(self.n_items))
            st["last_user"].fill_(-st["step"]); st["last_item"].fill_(-st["step"])    # up to date, nothing pending
        S = _cabi.LazyAdamState()
        for names, tensors in ((("theta", "beta", "xi", "eta"), params), (("m_theta", "m_beta", "m_xi", "m_eta"), st["m"]),
                               (("v_theta", "v_beta", "v_xi", "v_eta"), st["v"]), (("g_theta", "g_beta", "g_xi", "g_eta"), grads)):
            for nm, t in zip(names, tensors):
                setattr(S, nm, t.data_ptr())
        for nm, t in (("last_user", st["last_user"]), ("last_item", st["last_item"]), ("claim_user", st["claim_user"]),
                      ("claim_item", st["claim_item"]), ("step_size", tab[0]), ("bc2_sqrt", tab[1])):
            setattr(S, nm, t.data_ptr())
        S.beta1, S.beta2, S.eps = beta1, beta2, eps
        S.tail1, S.tail2, S.pow5 = (tails[0].data_ptr(), tails[1].data_ptr(), tails[2].data_ptr()) if closed_form else (None, None, None)
        S.n_pow = total + 2
        id_bytes = 8 if u_all.dtype == torch.int64 else 4
        acc = torch.zeros(epochs, dtype=torch.float64, device=dev)      # one loss per epoch, read back once at the end
        losses = []
        with torch.cuda.device(dev), torch.no_grad():
            # no host synchronisation inside the loop, and the permutations come from host threads running ahead
            perms = iter(_EpochPermutations(n, epochs, dev, prefetch=on_epoch is None)) if shuffle else None
            for ep in range(epochs):
                if shuffle:
                    perm = next(perms)
                    if ep == 0:
                        self._ev0.record(torch.cuda.current_stream(dev))
                    u_ep, i_ep, r_ep = u_all[perm].contiguous(), i_all[perm].contiguous(), r_all[perm].contiguous()
                else:
                    if ep == 0:
                        self._ev0.record(torch.cuda.current_stream(dev))
                    u_ep, i_ep, r_ep = u_all, i_all, r_all
                _cabi.call("pmf_hpf_map_lazy_epoch", C.byref(S), u_ep.data_ptr(), i_ep.data_ptr(), id_bytes, r_ep.data_ptr(), n,
                           batch_size, st["step"], self.user_scale.data_ptr(), self.item_scale.data_ptr(), self.n_users,
                           self.n_items, self.K, cfg.a, cfg.a_prime, cfg.b_prime, cfg.c, cfg.c_prime, cfg.d_prime,
                           acc[ep:].data_ptr(), self._bad.data_ptr(), _cabi.stream_ptr())
                st["step"] += steps_per_epoch
                if on_epoch is not None:
                    self._lazy_flush(S, st)
                    losses.append(float(acc[ep].item()))
                    on_epoch(ep, losses[-1])
            self._lazy_flush(S, st)
            if on_epoch is None:
                losses = [float(v) for v in acc.cpu().tolist()]
        self.check_ids()
        return losses

    def _lazy_flush(self, S, st):
        import ctypes as C
        _cabi.call("pmf_hpf_map_lazy_flush", C.byref(S), self.n_users, self.n_items, self.K, st["step"], _cabi.stream_ptr())
