"""Row (b): the reference's import paths and config contracts, checked on CPU.  Where the reference checkout is
available (build container) the config dataclasses are compared field by field with the reference's own."""
import ast
import dataclasses
import importlib.util
import os
import sys

import pytest

REF = os.environ.get("PMF_REFERENCE_ROOT", "/root/reference")
PAIRS = [("poisson_mf_cavi", "PoissonMFCAVIConfig", "PoissonMFCAVI"), ("hpf_cavi", "HPF_CAVI_Config", "HPF_CAVI"),
         ("poisson_mf_extended_cavi", "PoissonMFExtendedCAVIConfig", "PoissonMFExtendedCAVI"),
         ("gaussian_mf_cavi", "GaussianMFCAVIConfig", "GaussianMFCAVI"),
         ("gaussian_mf_cavi_bias", "GaussianMFCAVIConfig", "GaussianMFCAVI"),
         ("hpf_pytorch", "HPF_PyTorch_Config", "HPF_PyTorch")]


def test_install_routes_reference_import_paths():
    from prob_matrix_factorization_b200 import dropin
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    try:
        dropin.install()
        from src.models.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig  # noqa: F401
        from src.models.hpf_cavi import HPF_CAVI, HPF_CAVI_Config  # noqa: F401
        from src.models.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig  # noqa: F401
        from src.models.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config  # noqa: F401
        from src.models.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig  # noqa: F401
        from src.evaluation.metrics import rmse, macro_mae  # noqa: F401
        assert PoissonMFCAVI.__module__.startswith("prob_matrix_factorization_b200")
        assert HPF_PyTorch.__module__.startswith("prob_matrix_factorization_b200")
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("mod,cfg,cls", PAIRS)
def test_config_round_trip(mod, cfg, cls):
    """asdict -> str -> literal_eval -> Config(**d) (tune_all_models.py:314-317, compare_models.py:41,70)."""
    m = importlib.import_module(f"prob_matrix_factorization_b200.{mod}")
    C = getattr(m, cfg)
    c = C()
    d = ast.literal_eval(str(dataclasses.asdict(c)))
    assert C(**d) == c
    assert set(C.__annotations__) == {f.name for f in dataclasses.fields(C)}


def test_best_hyperparams_file_lines_construct_configs():
    """The dict literals of the reference's best_hyperparams.txt:3-6 must construct our configs."""
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVIConfig
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI_Config
    from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch_Config
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVIConfig
    GaussianMFCAVIConfig(**{'n_factors': 30, 'sigma2': 0.3, 'eta_theta2': 0.5, 'eta_beta2': 0.5, 'eta_bias2': 1.0,
                            'max_iter': 100, 'tol': 0.001, 'random_state': 42, 'verbose': True})
    PoissonMFCAVIConfig(**{'n_factors': 40, 'a0': 0.1, 'b0': 0.5, 'max_iter': 150, 'tol': None, 'random_state': 42, 'verbose': True})
    HPF_CAVI_Config(**{'n_factors': 20, 'a': 0.3, 'a_prime': 5.0, 'b_prime': 5.0, 'c': 0.3, 'c_prime': 5.0, 'd_prime': 5.0,
                       'max_iter': 100, 'tol': None, 'random_state': 42, 'verbose': True})
    raw = {'n_factors': 10, 'a': 1.0, 'a_prime': 1.0, 'b_prime': 1.0, 'c': 1.0, 'c_prime': 1.0, 'd_prime': 1.0, 'lr': 0.0005,
           'batch_size': 1024, 'epochs': 50, 'device': 'cpu', 'verbose': True}
    valid = HPF_PyTorch_Config.__annotations__.keys()                      # compare_models.py:265-270
    HPF_PyTorch_Config(**{k: v for k, v in raw.items() if k in valid})


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "models")), reason="reference checkout not present")
@pytest.mark.parametrize("mod,cfg,cls", PAIRS)
def test_configs_and_methods_match_reference(mod, cfg, cls):
    spec = importlib.util.spec_from_file_location(f"_ref_{mod}", os.path.join(REF, "src", "models", mod + ".py"))
    ref = importlib.util.module_from_spec(spec)
    sys.path.insert(0, REF)
    try:
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
    ours = importlib.import_module(f"prob_matrix_factorization_b200.{mod}")
    rf = [(f.name, f.default) for f in dataclasses.fields(getattr(ref, cfg))]
    of = [(f.name, f.default) for f in dataclasses.fields(getattr(ours, cfg))]
    assert rf == of
    ref_methods = {n for n in vars(getattr(ref, cls)) if not n.startswith("_") and callable(getattr(getattr(ref, cls), n))}
    our_cls = getattr(ours, cls)
    missing = {n for n in ref_methods if not hasattr(our_cls, n)}
    assert not missing, missing


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/experiments"), reason="reference checkout not present")
def test_reference_scripts_import_the_engine_through_dropin():
    """The reference's own driver modules (train_*_full.py, compare_models.py, tune_all_models.py), imported UNMODIFIED
    from the read-only checkout with ``dropin.install`` + ``stub_matplotlib``, bind our classes (no fit here: no GPU)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.dont_write_bytecode = True\n"
        "from prob_matrix_factorization_b200 import dropin\n"
        "dropin.stub_matplotlib(); dropin.install('/root/reference')\n"
        "import src.experiments.train_poisson_full as tp, src.experiments.train_hpf_cavi_full as th\n"
        "import src.experiments.train_gaussian_full as tg, src.experiments.compare_models as cm\n"
        "import prob_matrix_factorization_b200.poisson_mf_cavi as ours_p, prob_matrix_factorization_b200.hpf_cavi as ours_h\n"
        "import prob_matrix_factorization_b200.gaussian_mf_cavi_bias as ours_g, prob_matrix_factorization_b200.hpf_pytorch as ours_t\n"
        "assert tp.PoissonMFCAVI is ours_p.PoissonMFCAVI and tp.PoissonMFCAVIConfig is ours_p.PoissonMFCAVIConfig\n"
        "assert th.HPF_CAVI is ours_h.HPF_CAVI and tg.GaussianMFCAVI is ours_g.GaussianMFCAVI\n"
        "assert cm.HPF_PyTorch is ours_t.HPF_PyTorch and cm.PoissonMFCAVI is ours_p.PoissonMFCAVI\n"
        "hp = cm.load_best_hyperparams('/root/reference/best_hyperparams.txt')\n"
        "cfg = ours_p.PoissonMFCAVIConfig(**hp['PoissonMF']); assert cfg.n_factors > 0\n"
        "print('DROPIN_IMPORTS_OK')\n")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=repo)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert "DROPIN_IMPORTS_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
