#!/usr/bin/env bash
# Installs the UNMODIFIED reference files the reference arms need into baseline/_ref/ (git-ignored: no reference
# source enters the history; NOT gpurun-ignored: the directory travels to the GPU box with the working tree).
#   tools/install_ref.sh [/root/reference]
# The reference has no build step (setup.py is a bare stub naming one package whose __init__ creates directories),
# so "install" = copy the files of the hot path and of cfg5's model, byte for byte, keeping their relative paths.
# baseline/ref_loader.py imports them by path with the two stub modules (matplotlib, deepclustering2) that the
# arithmetic never touches.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="${HERE}/baseline/_ref"
FILES=(
  contrastyou/losses/contrast_loss3.py      # SelfPacedSupConLoss / SupConLoss1: the reference arm of bench.py
  contrastyou/projectors/heads.py           # ProjectionHead / DenseProjectionHead (cfg5, decoder harness)
  contrastyou/projectors/nn.py
  semi_seg/arch/unet.py                     # UNet (cfg5)
  semi_seg/arch/utils.py
)
if [ ! -d "${SRC}" ]; then
  echo "install_ref: ${SRC} not found (GPU box?): keeping ${DST} as it is" >&2
  exit 0
fi
for f in "${FILES[@]}"; do
  mkdir -p "${DST}/$(dirname "$f")"
  cp "${SRC}/$f" "${DST}/$f"
done
( cd "${DST}" && sha256sum "${FILES[@]}" > MANIFEST.sha256 )
echo "installed ${#FILES[@]} reference files into ${DST}"
