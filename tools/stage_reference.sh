#!/bin/sh
# Stage the reference's own implementation of the hot path for the timed baseline (bench.py --impl reference and the
# `cpu_baseline` / `reference_on_gpu` legs): copies /root/reference/model.py -- the one file the path lives in
# (CFM.forward :1136, BASECFM Euler loop :1084-1109, Decoder :834-1048; needs only torch + einops + numpy) -- into the
# git-ignored baseline/_ref/, which travels to the GPU box with the gpurun snapshot.  No reference source enters the
# history.  /root/reference does not exist on the GPU box; nothing reads it at run time.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
if [ ! -f "$SRC/model.py" ]; then
  echo "stage_reference: $SRC/model.py not found (nothing staged)" >&2
  exit 0
fi
mkdir -p "$ROOT/baseline/_ref"
cp -f "$SRC/model.py" "$ROOT/baseline/_ref/model.py"
chmod u+w "$ROOT/baseline/_ref/model.py"
echo "staged $SRC/model.py -> baseline/_ref/model.py"
# the vocoder row: the vendored HiFi-GAN package the reference calls after the mel (main.py:134-150)
if [ -d "$SRC/hifigan" ]; then
  mkdir -p "$ROOT/baseline/_ref/hifigan"
  for f in __init__.py models.py config.py env.py xutils.py denoiser.py; do
    cp -f "$SRC/hifigan/$f" "$ROOT/baseline/_ref/hifigan/$f"
    chmod u+w "$ROOT/baseline/_ref/hifigan/$f"
  done
  echo "staged $SRC/hifigan/{models,config,env,xutils,denoiser}.py -> baseline/_ref/hifigan/"
fi
