"""Make the reference's import paths resolve to the B200 engine.

The reference's tuning / compare / train scripts import the models by module path
(compare_models.py:17-20, tune_all_models.py:10-14, train_*_full.py:9-10):

    from src.models.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    from src.models.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    from src.models.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    from src.models.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config
    from src.models.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig   (run_poisson_mf_extended.py:4)

``install()`` registers this package's modules under those names in ``sys.modules`` so that the scripts
run unchanged on top of libpmf_b200:

    python -m prob_matrix_factorization_b200.dropin --reference /path/to/reference src.experiments.train_all_models --dataset_mode full
    python -m prob_matrix_factorization_b200.dropin --stub-matplotlib --script my_driver.py [args]     (a file instead of a module)

If the reference checkout is on ``sys.path`` its other packages (``src.experiments``, ``src.data`` ...)
are used as they are; without it, stub ``src`` / ``src.models`` / ``src.evaluation`` packages are created so
code written against the reference's import paths still works.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types

MODEL_MODULES = ("poisson_mf_cavi", "poisson_mf_extended_cavi", "hpf_cavi", "gaussian_mf_cavi", "gaussian_mf_cavi_bias",
                 "hpf_pytorch")


def _package(name, path=None):
    mod = sys.modules.get(name)
    if mod is None:
        try:
            mod = importlib.import_module(name)
        except ImportError:
            mod = types.ModuleType(name)
            mod.__path__ = [] if path is None else [path]
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(_package(parent), child, mod)
    return mod


def install(reference_root=None):
    """Route ``src.models.<model>`` (and ``src.evaluation.metrics`` if absent) to this engine."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _package("src")
    models = _package("src.models")
    for short in MODEL_MODULES:
        ours = importlib.import_module(f"{__package__}.{short}")
        sys.modules[f"src.models.{short}"] = ours
        setattr(models, short, ours)
    try:
        importlib.import_module("src.evaluation.metrics")
    except ImportError:
        ev = _package("src.evaluation")
        ours = importlib.import_module(f"{__package__}.metrics")
        sys.modules["src.evaluation.metrics"] = ours
        setattr(ev, "metrics", ours)
    return [f"src.models.{s}" for s in MODEL_MODULES]


def stub_matplotlib():
    """Plots are out of scope; let scripts that ``import matplotlib.pyplot`` run where it is not installed."""
    try:
        import matplotlib  # noqa: F401
        return False
    except ImportError:
        pass

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__") and name.endswith("__"):      # inspect / importlib probe modules for __file__, __spec__ ...
                raise AttributeError(name)
            return _Anything(name)

        def __call__(self, *a, **k):
            return _Anything("call")

        def __iter__(self):
            return iter(())

    root = _Anything("matplotlib")
    root.__path__ = []
    sys.modules["matplotlib"] = root
    sys.modules["matplotlib.pyplot"] = _Anything("matplotlib.pyplot")
    return True


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ref = os.environ.get("PMF_REFERENCE_ROOT")
    script = None
    while argv and argv[0].startswith("--"):
        flag = argv.pop(0)
        if flag == "--reference":
            ref = argv.pop(0)
        elif flag == "--stub-matplotlib":
            stub_matplotlib()
        elif flag == "--script":
            script = argv[0]
            break
        else:
            raise SystemExit(f"unknown flag {flag}")
    if not argv:
        raise SystemExit(__doc__)
    install(ref)
    sys.argv = argv
    if script is not None:
        runpy.run_path(script, run_name="__main__")
    else:
        runpy.run_module(argv[0], run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
