#!/bin/bash
# Builds the library of another revision next to the in-tree one, for same-box A/B runs:
#   tools/build_alt.sh HEAD~1 before     -> build/alt/before.so
#   gpurun -- 'bash tools/gpu_ab_builds.sh TAG build/alt/before.so default'
# (build/ is git-ignored but travels to the GPU box; the A/B pairs r04b / r04d / r05b under
# profiles/ were taken this way, interleaved twice)
set -e
REV=${1:?revision}; NAME=${2:?name}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/build/src_$NAME
rm -rf "$SRC"; mkdir -p "$SRC" "$ROOT/build/alt"
git -C "$ROOT" archive "$REV" jaxsplat_b200/csrc include | tar -x -C "$SRC"
make -C "$SRC/jaxsplat_b200/csrc" -j8 OBJDIR="$ROOT/build/obj_$NAME" OUT="$ROOT/build/alt/$NAME.so" \
     PYMOD="$ROOT/build/alt/_unused.so" "$ROOT/build/alt/$NAME.so" 2>&1 | grep -E "error|warning" || true
ls -la "$ROOT/build/alt/$NAME.so"
