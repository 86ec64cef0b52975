"""B200-native (sm_100a) implementation of UCF-VIT's transformer-block training hot path behind the reference's
module API (same sub-package layout as `UCF_VIT`: simple/, fsdp/, dataloaders/, ddpm/, utils/)."""


def install_as(name: str = "UCF_VIT"):
    """Make `import UCF_VIT.simple.arch`, `from UCF_VIT.utils.misc import configure_optimizer`, ... resolve to this
    package, so a reference training script runs on these kernels with no edit beyond one line before its imports:

        import ucf_vit_b200; ucf_vit_b200.install_as("UCF_VIT")

    Sub-modules of the reference that lie outside the hot path and are not provided here (datasets/, datamodule,
    inference helpers -- INTEGRATION.md lists them) still import from the reference if it is installed."""
    import importlib
    import pkgutil
    import sys
    pkg = sys.modules[__name__]
    sys.modules[name] = pkg
    for m in pkgutil.walk_packages(pkg.__path__, __name__ + "."):
        short = m.name[len(__name__):]
        if short.split(".")[-1].startswith("_"):
            continue
        sys.modules[name + short] = importlib.import_module(m.name)
    return pkg
