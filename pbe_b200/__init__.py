"""pbe_b200 — B200-native (sm_100a) denoising hot path for Paint-by-Example (zhanwenchen/pbe)."""
__version__ = "0.1.0"


def install():
    """Make the reference's ``instantiate_from_config`` / sampler imports resolve to this package (see dropin.py)."""
    from .dropin import install as _install
    return _install()


def uninstall():
    from .dropin import uninstall as _uninstall
    return _uninstall()
