#!/bin/bash
# Stages the UNMODIFIED reference (pure Python: flow2gan/ + its wav<->mel fixtures) under
# baseline/_ref/ so that `bench.py --impl reference` can run the reference's own code path
# (flow2gan.get_model(checkpoint=...) -> model.infer, flow2gan/__init__.py:29-47,
# flow2gan/models/generator.py:327-366) on the GPU box's host cores.  baseline/_ref/ is
# git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored,
# so it travels to the box with the snapshot.  The reference has no setup.py / pyproject, so
# `pip install --target baseline/_ref /root/reference` has nothing to build: a plain copy IS the install.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -d "$SRC/flow2gan" ] || { echo "stage_reference: $SRC/flow2gan not found (GPU box?) -- keeping $DST as is"; exit 0; }
mkdir -p "$DST"
rm -rf "$DST/flow2gan" "$DST/test_data"
cp -r "$SRC/flow2gan" "$DST/flow2gan"
mkdir -p "$DST/test_data"
cp -r "$SRC/test_data/mel" "$SRC/test_data/wav" "$DST/test_data/" 2>/dev/null || true
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find flow2gan -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST/flow2gan" -name '*.py' | wc -l) reference files under $DST"
