#!/usr/bin/env bash
# Builds the REAL reference (sources read in place from $HANA_REFERENCE, default
# /root/reference/Hana-SoftwareRenderer) into oracle/_ref/:
#   libhana_ref.so       plain -O2 build, for timing the CPU baseline
#   libhana_ref_inst.so  instrumented build (prim-ID buffer, counters, stage hooks)
#   assets/*.hscene      the bundled scenes re-packed (a2v stream + decoded textures)
# Nothing from the reference is written into the repository: graphics.cpp is
# patched on the fly (sed -> compiler stdin); objects live in a temp dir;
# oracle/_ref/ is git-ignored (it still travels to the GPU box with gpurun).
#
# Accommodations: shim/Color.h (D3), forced
# includes (D4), --allow-multiple-definition (D2), v2fs[3] -> v2fs[10] (D1).
# No -march/-mfma/-ffast-math: x86-64 SSE2 float, no FMA contraction.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${HANA_REFERENCE:-/root/reference/Hana-SoftwareRenderer}"
OUT="$HERE/_ref"
if [ ! -f "$R/graphics.cpp" ]; then echo "reference not found at $R" >&2; exit 3; fi
TMP="$(mktemp -d /tmp/hana_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
# NOT $CXX: this image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ statically,
# which crashes once another libstdc++ user (numpy, torch) is loaded in the same process.
CXX="${HANA_CXX:-g++}"
F="-std=c++17 -O2 -fPIC -ffp-contract=off -w -I$R -I$HERE/shim -include cstring -include limits -include algorithm -include cstdio"
for f in vector color maths model IShader renderbuffer tgaimage gameobject camera scene; do
  $CXX $F -c "$R/$f.cpp" -o "$TMP/$f.o" &
done
wait
# D1 only (timing build)
LC_ALL=C sed 's/shader_struct_v2f v2fs\[3\];/shader_struct_v2f v2fs[10];/' "$R/graphics.cpp" \
  | $CXX $F -x c++ -c - -o "$TMP/graphics_plain.o"
# D1 + hooks (instrumented build); the hook wrappers are appended to the stream
{ LC_ALL=C sed \
    -e 's/shader_struct_v2f v2fs\[3\];/shader_struct_v2f v2fs[10];/' \
    -e 's/v2fs\[j\] = draw_data->shader->vertex(&a2v);/href_hook_a2v(i, j, \&a2v); v2fs[j] = draw_data->shader->vertex(\&a2v);/' \
    -e 's/^\t\t\trasterize_triangle(draw_data, ret_v2fs);/\t\t\thref_hook_prim(i, j); rasterize_triangle(draw_data, ret_v2fs);/' \
    -e 's/^\t\t\t\trender_buffer->set_depth(P.x, P.y, frag_depth);/\t\t\t\trender_buffer->set_depth(P.x, P.y, frag_depth); href_hook_frag(P.x, P.y);/' \
    -e 's/^\t\t\tVector3f barycentric_weights = barycentric(screen_coords\[0\], screen_coords\[1\], screen_coords\[2\], P);/&href_hook_bbox();/' \
    -e 's/^\t\t\tif (barycentric_weights.x < 0 || barycentric_weights.y < 0 || barycentric_weights.z < 0) continue;/& href_hook_inside();/' \
    "$R/graphics.cpp"; cat "$HERE/ref_graphics_hooks.inc"; } \
  | $CXX $F -include "$HERE/ref_hooks.h" -x c++ -c - -o "$TMP/graphics_inst.o"
$CXX $F -c "$HERE/ref_driver.cpp" -o "$TMP/driver_plain.o"
$CXX $F -DHREF_INSTRUMENTED -c "$HERE/ref_driver.cpp" -o "$TMP/driver_inst.o"
COMMON="$TMP/vector.o $TMP/color.o $TMP/maths.o $TMP/model.o $TMP/IShader.o $TMP/renderbuffer.o $TMP/tgaimage.o $TMP/gameobject.o $TMP/camera.o $TMP/scene.o"
# the plain build still needs the hook symbols referenced by nothing -> none needed
$CXX -shared -o "$OUT/libhana_ref.so" $COMMON "$TMP/graphics_plain.o" "$TMP/driver_plain.o" -Wl,--allow-multiple-definition -Wl,-Bsymbolic
$CXX -shared -o "$OUT/libhana_ref_inst.so" $COMMON "$TMP/graphics_inst.o" "$TMP/driver_inst.o" -Wl,--allow-multiple-definition -Wl,-Bsymbolic
# The drop-in test build: the reference's scene code + this repo's graphics_draw_triangle shim INSTEAD of
# graphics.cpp, linked against libhana_b200.so (needs a GPU to run; only built when the library exists).
PKG="$HERE/../hana-softwarerenderer_b200"
if [ -f "$PKG/libhana_b200.so" ]; then
  $CXX $F -c "$PKG/host/graphics_dropin.cpp" -o "$TMP/graphics_dropin.o"
  $CXX -shared -o "$OUT/libhana_ref_dropin.so" $COMMON "$TMP/graphics_dropin.o" "$TMP/driver_plain.o" \
      -Wl,--allow-multiple-definition -Wl,-Bsymbolic -L"$PKG" -lhana_b200 -Wl,-rpath,'$ORIGIN/../../hana-softwarerenderer_b200'
  echo "built $OUT/libhana_ref_dropin.so"
fi
echo "built $OUT/libhana_ref.so $OUT/libhana_ref_inst.so"
