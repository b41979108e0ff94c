#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the C++ host side, no GPU needed:
#   sanitize_frontend: RON reader, formula evaluator, scene evaluation, animation player on every given scene file and on
#                      30 random mutations of each;
#   sanitize_rewriter: the lexical GLSL -> CUDA rewriter on every snippet of every scene and on 11 mutations of each;
#   sanitize_png:      the PNG codec -- round trips, every prefix, thousands of mutations with repaired CRCs (PNG_FILES="a.png b.png"
#                      adds real files).
# usage: tools/sanitize_host.sh tests/fixtures/two_spheres.ron /root/reference/scenes/*.ron
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${TMPDIR:-/tmp}/portal_b200_sanitize"
mkdir -p "$OUT"
python -m portal_b200.build > /dev/null      # makes sure the embedded device sources (pe_device_src.cpp) exist
FLAGS="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I$ROOT/portal_b200/csrc -I$ROOT/portal_b200/csrc/host -I$ROOT/include"
H="$ROOT/portal_b200/csrc/host"
g++ $FLAGS "$ROOT/tests/host_harness/sanitize_frontend.cpp" "$H/ph_ron.cpp" "$H/ph_formula.cpp" "$H/ph_scene.cpp" "$H/ph_matsqrt.cpp" "$H/ph_anim.cpp" -o "$OUT/frontend"
g++ $FLAGS "$ROOT/tests/host_harness/sanitize_rewriter.cpp" "$ROOT/portal_b200/csrc/pe_codegen.cpp" "$ROOT/portal_b200/csrc/pe_device_src.cpp" \
    "$H/ph_ron.cpp" "$H/ph_formula.cpp" "$H/ph_scene.cpp" "$H/ph_matsqrt.cpp" -o "$OUT/rewriter"
g++ $FLAGS "$ROOT/tests/host_harness/sanitize_png.cpp" "$H/ph_png.cpp" -o "$OUT/png"
"$OUT/png" $PNG_FILES
"$OUT/frontend" "$@"
"$OUT/rewriter" "$@"
echo "sanitizers: no reports"
