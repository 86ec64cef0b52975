#!/bin/bash
# Installs the UNMODIFIED reference (irlyngaas/UCF-VIT) into the git-ignored baseline/_ref/ so that
# `bench.py --impl reference` can time the reference's own modules on the host cores.  Offline: the wheelhouse has none
# of its dependencies (monai, timm, xformers, torchdata, opencv), so only the package itself is installed (--no-deps);
# the missing third-party imports are satisfied at run time by the stand-ins under oracle/shims/.
# /root/reference is read-only, so pip builds from a copy under /tmp.  baseline/_ref travels to the GPU box with gpurun.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d /root/reference ] || { echo "[install_ref] /root/reference absent: keeping $(ls "$HERE/_ref" 2>/dev/null | wc -l) prebuilt entries"; exit 0; }
rm -rf /tmp/ucf_ref_copy "$HERE/_ref"
cp -r /root/reference /tmp/ucf_ref_copy
python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$HERE/_ref" /tmp/ucf_ref_copy
rm -rf /tmp/ucf_ref_copy
echo "[install_ref] installed: $(ls "$HERE/_ref")"
