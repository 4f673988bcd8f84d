"""Reading the reference's snapshots.

The reference saves WHOLE-MODEL pickles (``torch.save(model, path)`` through ml_utilities' SnapshotManager,
setup_functions.py:134-164; read back with ``load_to_cpu``, audio_model.py:287-290).  Such a file names the
reference's own modules (``audio_model``, ``scalogram_model``, ``constant_q_transform`` ...), which are not importable
next to this package -- and its conv layers are stock ``torch.nn`` modules.  Because every module here keeps the
reference's attribute names, the *state_dict* of a snapshot maps one to one onto a model built by
``configs.setup_model``; this module extracts it without importing any reference code:

    model, pre, _ = cpc_b200.configs.setup_model(...)          # same experiment dicts as the reference run
    step = cpc_b200.snapshots.load_reference_snapshot(model, "snapshots/e24_120000")

Reference classes inside the pickle are materialised as inert stand-ins (``nn.Module`` subclasses for modules, plain
objects otherwise).  Every other global must be on an exact allow-list (tensor / storage rebuild functions, dtypes,
``torch.nn.modules.*`` layer classes, containers, numpy array reconstruction); anything else -- including callables that
are merely reachable under the ``torch`` package -- is refused.
"""
import collections
import io
import pickle
import re

import torch
import torch.nn as nn

# modules of the reference whose classes may appear in a snapshot
REFERENCE_MODULES = ("audio_model", "scalogram_model", "constant_q_transform", "attention_model", "transformer",
                     "contrastive_estimation_training", "audio_dataset", "classification_model", "ml_utilities")

_stub_cache = {}


class _InertObject:
    """Stand-in for a non-module reference object (e.g. an ActivationRegister referenced by a writer)."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def _stub_class(module, name, is_module=True):
    key = (module, name)
    if key not in _stub_cache:
        base = nn.Module if is_module else _InertObject
        _stub_cache[key] = type(name, (base,), {"__module__": "cpc_b200.snapshots.<%s>" % module,
                                                "forward": lambda self, *a, **k: (_ for _ in ()).throw(
                                                    RuntimeError("snapshot stand-in: build the model with "
                                                                 "cpc_b200.configs.setup_model and load the state_dict"))})
    return _stub_cache[key]


# reference classes that are NOT nn.Modules
_NON_MODULES = {"ActivationRegister"}


# Everything else a model pickle legitimately needs, as EXACT globals (a pickle can name any importable callable, and
# re-exports such as torch.serialization.os or torch.hub.load live under the torch root, so roots are not enough).
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("collections", "deque"),
    ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._utils", "_rebuild_qtensor"),
    ("torch._utils", "_rebuild_device_tensor_from_numpy"), ("torch.storage", "_load_from_bytes"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"), ("torch.nn.parameter", "Parameter"),
    ("torch.nn.parameter", "Buffer"), ("torch._tensor", "_rebuild_from_type_v2"), ("torch", "Tensor"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("_codecs", "encode"), ("copyreg", "_reconstructor"),
}
_ALLOWED_BUILTINS = {"set", "frozenset", "dict", "list", "tuple", "slice", "range", "complex", "bytearray", "object",
                     "int", "float", "bool", "str", "bytes"}
_TORCH_STORAGES = re.compile(r"^(Untyped|Float|Double|Half|BFloat16|Long|Int|Short|Char|Byte|Bool|Complex(Float|Double))Storage$")
_TORCH_DTYPES = re.compile(r"^(float(16|32|64)|bfloat16|half|float|double|u?int(8|16|32|64)|long|int|short|bool|complex(64|128))$")


class _SnapshotUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if root in REFERENCE_MODULES or root == "__main__":
            return _stub_class(module, name, is_module=name not in _NON_MODULES)
        allowed = ((module, name) in _ALLOWED_GLOBALS
                   or (module in ("builtins", "__builtin__") and name in _ALLOWED_BUILTINS)
                   or (module == "torch" and (_TORCH_STORAGES.match(name) or _TORCH_DTYPES.match(name))))
        if not allowed and module.startswith("torch.nn.modules.") and "." not in name:
            candidate = super().find_class(module, name)         # stock layers inside the model: nn.Module classes only
            allowed = isinstance(candidate, type) and issubclass(candidate, nn.Module)
        if allowed:
            found = super().find_class(module, name)
            if not isinstance(found, type(re)):                  # never hand out a module object
                return found
        raise pickle.UnpicklingError("snapshot names %s.%s, which is neither a reference class nor one of the torch / numpy "
                                     "/ container globals a model pickle needs" % (module, name))


class _PickleModule:
    """The ``pickle_module`` torch.load expects (Unpickler + load), with the remapping unpickler."""
    __name__ = "cpc_b200.snapshots"
    Unpickler = _SnapshotUnpickler

    @staticmethod
    def load(file, **kwargs):
        return _SnapshotUnpickler(file, **kwargs).load()


def read_reference_snapshot(path):
    """The object graph of a reference snapshot with every reference class replaced by an inert stand-in (CPU tensors)."""
    return torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)


def extract_state_dict(path):
    """``state_dict`` of the model stored in a reference snapshot; keys are exactly the reference's."""
    obj = read_reference_snapshot(path)
    if isinstance(obj, dict):                                   # a plain state_dict (or {'state_dict': ...}) is fine too
        inner = obj.get("state_dict", obj)
        return collections.OrderedDict((k, v) for k, v in inner.items() if torch.is_tensor(v))
    if isinstance(obj, nn.DataParallel) or hasattr(obj, "module") and isinstance(getattr(obj, "module"), nn.Module):
        obj = obj.module                                        # setup_functions.py:112-115 wraps multi-GPU models
    if not isinstance(obj, nn.Module):
        raise TypeError("snapshot holds a %s, not a model" % type(obj).__name__)
    return obj.state_dict()


def snapshot_step(path):
    """Training step encoded in a snapshot name ``<name>_<step>`` (setup_functions.py:156), or None."""
    m = re.search(r"_(\d+)$", str(path).rstrip("/").split("/")[-1].split(".")[0])
    return int(m.group(1)) if m else None


def load_reference_snapshot(model, path, strict=True):
    """Load the parameters and buffers of a reference snapshot into ``model`` (built by ``configs.setup_model`` with
    the same experiment).  Returns the training step encoded in the file name (None if absent)."""
    state = extract_state_dict(path)
    target = model.module if isinstance(model, nn.DataParallel) else model
    # a DataParallel-wrapped reference model prefixes its keys with 'module.'
    if state and all(k.startswith("module.") for k in state) and not any(k.startswith("module.") for k in target.state_dict()):
        state = collections.OrderedDict((k[len("module."):], v) for k, v in state.items())
    target.load_state_dict(state, strict=strict)
    return snapshot_step(path)


def save_snapshot(model, path):
    """Write ``model``'s state_dict (the portable form; a whole-model pickle of THIS package would name cpc_b200's
    classes and could not be read by the reference either)."""
    target = model.module if isinstance(model, nn.DataParallel) else model
    buffer = io.BytesIO()
    torch.save(collections.OrderedDict((k, v.detach().cpu()) for k, v in target.state_dict().items()), buffer)
    with open(path, "wb") as fh:
        fh.write(buffer.getvalue())
