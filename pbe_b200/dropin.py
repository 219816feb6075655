"""The drop-in switch: ``pbe_b200.install()`` makes the reference's own plugin mechanism resolve to this package.

The reference builds its model with ``instantiate_from_config`` (``ldm/util.py:78-93``: ``importlib.import_module`` of a
dotted ``target:``) from ``configs/v1.yaml`` and imports its samplers by module path (``scripts/inference.py:17-19``).
``install()`` registers this package's mirrors in ``sys.modules`` under exactly those module paths, so the unedited yaml
and the unedited scripts pick them up:

=================================================  =========================================  ======================
reference module path                              what the reference takes from it           served by
=================================================  =========================================  ======================
``ldm.models.diffusion.ddpm``                      ``LatentDiffusion`` (configs/v1.yaml:3)     ``pbe_b200.diffusion``
``ldm.models.diffusion.latent_diffusion``          ``LatentDiffusion`` (where the fork moved it)  ``pbe_b200.diffusion``
``ldm.modules.diffusionmodules.openaimodel``       ``UNetModel`` (v1.yaml:31)                  ``pbe_b200.unet``
``ldm.models.autoencoder``                         ``AutoencoderKL`` (v1.yaml:49)              ``pbe_b200.vae``
``ldm.modules.encoders.modules``                   ``FrozenCLIPImageEmbedder`` (v1.yaml:72)    ``pbe_b200.clip``
``ldm.models.diffusion.plms`` / ``.ddim``          ``PLMSSampler`` / ``DDIMSampler``           ``pbe_b200.samplers``
=================================================  =========================================  ======================

Everything else of the reference (``ldm.util``, ``ldm.data``, the scripts) stays the reference's.  When the reference
tree is not importable at all the parent packages are created as empty namespace stubs, so ``from
ldm.models.diffusion.plms import PLMSSampler`` works standalone too.  ``uninstall()`` restores ``sys.modules``.
"""
from __future__ import annotations

import importlib
import sys
import types
from typing import Dict

_ALIASES = {
    "ldm.models.diffusion.ddpm": "pbe_b200.diffusion",
    "ldm.models.diffusion.latent_diffusion": "pbe_b200.diffusion",
    "ldm.modules.diffusionmodules.openaimodel": "pbe_b200.unet",
    "ldm.models.autoencoder": "pbe_b200.vae",
    "ldm.modules.encoders.modules": "pbe_b200.clip",
    "ldm.models.diffusion.plms": "pbe_b200.samplers",
    "ldm.models.diffusion.ddim": "pbe_b200.samplers",
}
_saved: Dict[str, object] = {}
_installed = False


def get_obj_from_str(string: str):
    """reference ``ldm/util.py:88-93``."""
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    """reference ``ldm/util.py:78-85`` (same sentinels, same KeyError); accepts dicts and OmegaConf nodes alike."""
    if "target" not in config:
        if config == "__is_first_stage__":
            return None
        elif config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", dict())
    return get_obj_from_str(config["target"])(**(params if params is not None else dict()))


def _ensure_parent_packages(name: str) -> None:
    """Make every parent of ``name`` importable: the reference's package if it is on sys.path, else an empty stub."""
    parts = name.split(".")
    for i in range(1, len(parts)):
        prefix = ".".join(parts[:i])
        if prefix in sys.modules:
            continue
        try:
            importlib.import_module(prefix)
        except Exception:
            stub = types.ModuleType(prefix)
            stub.__path__ = []          # a package, with nothing of its own to import
            stub.__pbe_b200_stub__ = True
            _saved.setdefault(prefix, None)
            sys.modules[prefix] = stub
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], stub)


def install() -> Dict[str, str]:
    """Register the mirrors under the reference's module paths (idempotent).  Returns {reference path: serving module}."""
    global _installed
    for ref_name, ours in _ALIASES.items():
        mod = importlib.import_module(ours)
        _ensure_parent_packages(ref_name)
        if not _installed:
            _saved.setdefault(ref_name, sys.modules.get(ref_name))
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        if parent in sys.modules:
            try:
                setattr(sys.modules[parent], leaf, mod)
            except Exception:
                pass
    _installed = True
    return dict(_ALIASES)


def uninstall() -> None:
    """Undo :func:`install`: the reference's own modules (if they had been imported) are put back."""
    global _installed
    for name, old in list(_saved.items()):
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
    _saved.clear()
    _installed = False


def installed() -> bool:
    return _installed
