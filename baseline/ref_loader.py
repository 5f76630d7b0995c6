"""Loader of the UNMODIFIED reference files installed under ``baseline/_ref/`` (``tools/install_ref.sh``).

REFERENCE ARM ONLY: ``bench.py`` (``--impl reference``, ``gpu_eager_reference``, ``cfg5``), ``tools/cfg5_step.py`` and the
tests that compare against the reference's own modules import this; the product package never does.

The files are byte-identical copies of ``/root/reference`` (checked against the committed ``REF_MANIFEST.sha256``).
They are imported by path under private package names, never through ``import contrastyou`` (its ``__init__`` creates
directories next to the package), with stub ``matplotlib`` / ``deepclustering2`` modules that the arithmetic of
``contrast_loss3.py`` never touches.
"""
from __future__ import annotations

import hashlib
import importlib
import pathlib
import sys
import types

HERE = pathlib.Path(__file__).resolve().parent
REF_ROOT = HERE / "_ref"
MANIFEST = HERE / "REF_MANIFEST.sha256"


class ReferenceUnavailable(RuntimeError):
    pass


def available() -> bool:
    return (REF_ROOT / "contrastyou" / "losses" / "contrast_loss3.py").exists()


def verify() -> None:
    """Every installed file must hash to the committed manifest (i.e. be the unmodified reference file)."""
    if not available():
        raise ReferenceUnavailable(f"{REF_ROOT} is empty: run tools/install_ref.sh in the build container")
    for line in MANIFEST.read_text().splitlines():
        digest, rel = line.split()
        got = hashlib.sha256((REF_ROOT / rel).read_bytes()).hexdigest()
        if got != digest:
            raise ReferenceUnavailable(f"baseline/_ref/{rel} differs from the reference file recorded in {MANIFEST.name}")


def _stubs() -> None:
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.get_backend = lambda: "agg"
            mpl.use = lambda *a, **k: None
            sys.modules["matplotlib"] = mpl
    for name in ("deepclustering2", "deepclustering2.configparser", "deepclustering2.configparser._utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["deepclustering2.configparser._utils"].get_config = lambda *a, **k: {}


def _package(name: str, directory: pathlib.Path):
    """A synthetic package whose submodules are the files of ``directory`` (relative imports inside them resolve)."""
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(directory)]
        sys.modules[name] = pkg
    return sys.modules[name]


_verified = False


def _load(pkg: str, directory: str, module: str):
    global _verified
    if not _verified:
        verify()
        _verified = True
    _stubs()
    from loguru import logger
    logger.disable(pkg)
    _package(pkg, REF_ROOT / directory)
    return importlib.import_module(f"{pkg}.{module}")


def loss_module():
    """``contrastyou/losses/contrast_loss3.py`` -> module with SelfPacedSupConLoss, SupConLoss1, is_normalized ..."""
    return _load("spcl_ref_losses", "contrastyou/losses", "contrast_loss3")


def unet_module():
    """``semi_seg/arch/unet.py`` -> module with UNet."""
    return _load("spcl_ref_arch", "semi_seg/arch", "unet")


def heads_module():
    """``contrastyou/projectors/heads.py`` -> module with ProjectionHead, DenseProjectionHead."""
    return _load("spcl_ref_projectors", "contrastyou/projectors", "heads")
