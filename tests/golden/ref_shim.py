"""Import shim for the REAL reference (/root/reference), usable only in the build container.

The reference needs `lightning` / `lightning_utilities` for four type aliases (utils/tensormask.py:4,
utils/helpers.py:6-9); they are not installed, so stub modules are registered before importing it.
Used by make_golden.py and by the optional `ref`-marked tests; never by the product path.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VGSLM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "speech"))


def import_reference():
    """returns (LVTR class, Hparams class, TensorMask class) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("lightning")
    stub("lightning.fabric")
    stub("lightning.fabric.utilities")
    stub("lightning.fabric.utilities.types", _DEVICE=object)
    stub("lightning.fabric.utilities.apply_func", _BLOCKING_DEVICE_TYPES=("cpu",), _TransferableDataType=object)
    stub("lightning_utilities")
    stub("lightning_utilities.core")
    stub("lightning_utilities.core.apply_func", apply_to_collection=lambda data, dtype, fn, *a, **k: fn(data))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from models.speech.lvtr import LVTR          # noqa: E402
    from hparams.hp import Hparams               # noqa: E402
    from utils.tensormask import TensorMask      # noqa: E402
    return LVTR, Hparams, TensorMask
