"""Copy the files of the read-only reference that its VAE-GSLM hot path needs into ``baseline/_ref/`` (git-ignored, travels
to the GPU box with the snapshot) so that ``bench.py --impl reference`` and ``bench.py``'s ``gpu_reference`` block can run
the UNMODIFIED reference modules there.  Called by ``__graft_entry__.build()`` in the build container, where
/root/reference exists; a no-op elsewhere.  Nothing under ``vae_gslm_b200/`` imports from ``baseline/``."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# what `from models.speech.lvtr import LVTR` pulls in (+ the sampler and the two configs)
TREES = ["hparams", "modules"]
FILES = ["__init__.py", "utils/__init__.py", "utils/tensormask.py", "utils/helpers.py", "utils/attr.py",
         "models/__init__.py", "models/speech/__init__.py", "models/speech/lvtr.py",
         "training_lib/__init__.py", "training_lib/losses.py",
         "trainers/__init__.py", "trainers/speech/__init__.py", "trainers/speech/sampler.py",
         "configs/train/speech/vae-gslm.yaml", "configs/infer/speech/vae-gslm.yaml"]


def vendor(src: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(src, "models", "speech")):
        return os.path.isdir(os.path.join(DEST, "models", "speech"))
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for t in TREES:
        shutil.copytree(os.path.join(src, t), os.path.join(DEST, t), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in FILES:
        s = os.path.join(src, f)
        d = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(s):
            shutil.copy2(s, d)
        elif f.endswith("__init__.py"):
            open(d, "w").close()                 # namespace packages in the reference: an empty marker is enough
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(DEST, "models", "speech"))


def load():
    """(LVTR, Hparams, TensorMask, masked_loss, config path) of the vendored reference, through the lightning shim of
    tests/golden/ref_shim.py (the reference needs four type aliases of a package that is not installed)."""
    import sys
    os.environ["VGSLM_REFERENCE_ROOT"] = DEST
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests", "golden"))
    import importlib
    import ref_shim
    importlib.reload(ref_shim)
    LVTR, Hparams, TensorMask = ref_shim.import_reference()
    from training_lib.losses import masked_loss
    return LVTR, Hparams, TensorMask, masked_loss, os.path.join(DEST, "configs", "train", "speech", "vae-gslm.yaml")


if __name__ == "__main__":
    print("vendored:", vendor())
