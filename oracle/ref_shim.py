"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (``/root/reference``) in this container so that
``oracle/make_golden.py`` can run it on CPU and freeze golden vectors under
``tests/golden/``.  The reference imports four packages that are not installed
here (``librosa``, ``mutagen``, ``ml_utilities``, ``matplotlib``); this module
registers in-memory stand-ins for them before the import.

The only stand-in that does arithmetic is ``librosa.filters.constant_q`` /
``librosa.time_frequency.cqt_frequencies`` (call sites
``constant_q_transform.py:108-112`` and ``:272-274``); it forwards to the
restatement in ``oracle/cpc_oracle.py`` (librosa <= 0.7 semantics, version not
pinned by the reference => that part of parity is UNPINNED).

``/root/reference`` does not exist on the GPU box, so nothing under ``tests -m gpu``,
``bench.py`` or ``smoke()`` may import this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CPC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(os.path.join(REFERENCE_ROOT, "audio_model.py"))


def _install_stubs():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import cpc_oracle

    if "librosa" not in sys.modules:
        lr = types.ModuleType("librosa")
        lr.filters = types.ModuleType("librosa.filters")
        lr.time_frequency = types.ModuleType("librosa.time_frequency")
        lr.core = types.ModuleType("librosa.core")

        def constant_q(sr, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0, window="hann",
                       filter_scale=1, pad_fft=True, norm=1, **kw):
            return cpc_oracle.constant_q_filters(sr, fmin, n_bins, bins_per_octave, filter_scale)

        def cqt_frequencies(n_bins, fmin, bins_per_octave=12, tuning=0.0):
            return cpc_oracle.cqt_frequencies(n_bins, fmin, bins_per_octave)

        lr.filters.constant_q = constant_q
        lr.time_frequency.cqt_frequencies = cqt_frequencies
        lr.cqt_frequencies = cqt_frequencies
        lr.load = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa.load stub"))
        sys.modules["librosa"] = lr
        sys.modules["librosa.filters"] = lr.filters
        sys.modules["librosa.time_frequency"] = lr.time_frequency
        sys.modules["librosa.core"] = lr.core

    for name in ("mutagen", "mutagen.mp3"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.MP3 = object
            sys.modules[name] = m
    sys.modules["mutagen"].mp3 = sys.modules["mutagen.mp3"]

    if "ml_utilities" not in sys.modules:
        mu = types.ModuleType("ml_utilities")
        tl = types.ModuleType("ml_utilities.train_logging")
        cu = types.ModuleType("ml_utilities.colab_utilities")
        pu = types.ModuleType("ml_utilities.pytorch_utilities")

        class AverageMeter:
            def __init__(self):
                self.reset()

            def reset(self):
                self.sum, self.count, self.avg, self.val = 0.0, 0, 0.0, 0.0

            def update(self, v, n=1):
                self.val = v
                self.sum += v * n
                self.count += n
                self.avg = self.sum / max(self.count, 1)

        class TensorboardLogger:
            def __init__(self, *a, **k):
                self.loss_meter = AverageMeter()
                self.score_meter = AverageMeter()

            def log(self, step):
                pass

        class _Dummy:
            def __init__(self, *a, **k):
                pass

        tl.TensorboardLogger = TensorboardLogger
        tl.AverageMeter = AverageMeter
        cu.GCSManager = _Dummy
        cu.SnapshotManager = _Dummy
        pu.parameter_count = lambda m: sum(p.numel() for p in m.parameters())
        def _any_symbol(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Dummy

        for mod in (tl, cu, pu):
            mod.__getattr__ = _any_symbol
        mu.train_logging, mu.colab_utilities, mu.pytorch_utilities = tl, cu, pu
        sys.modules["ml_utilities"] = mu
        sys.modules["ml_utilities.train_logging"] = tl
        sys.modules["ml_utilities.colab_utilities"] = cu
        sys.modules["ml_utilities.pytorch_utilities"] = pu

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        def _plt_symbol(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

        plt.__getattr__ = _plt_symbol
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_loaded = {}


def load_reference():
    """Import the reference's hot-path modules; returns a dict name -> module."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    for name in ("audio_model", "constant_q_transform", "scalogram_model", "attention_model",
                 "audio_dataset", "contrastive_estimation_training"):
        _loaded[name] = importlib.import_module(name)
    return _loaded


def load_reference_configs():
    load_reference()
    import importlib
    return importlib.import_module("configs.experiment_configs")
