#!/bin/sh
# build_app.sh <variant: k1|k31|kMD> -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED reference application from the sources where they lie under $REF (default /root/reference) with the
# reference's own flags (cellBP/Makefile.local_defs:10,17), the El Topo library included (eltopo3d/Makefile.inc), and links it twice:
#   oracle/_ref/cellBP_ref_<v>   main.cpp + El Topo + the reference's UL_growth_explicit.o          (the CPU reference, whole app)
#   oracle/_ref/cellBP_gpu_<v>   main.cpp + El Topo + wholecellfe_b200/csrc/ul_adapter.cpp + libwcfe  (the literal drop-in)
# Outside help is confined to oracle/app_shim.cpp (five BLAS/LAPACK symbols, srand/sprintf link wraps).  No reference source is copied
# or modified; object files and binaries go to oracle/_ref/ only.
set -e
V="$1"
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
R="$REF/wholeCellFE_$V"
OUT="$HERE/_ref"
OBJ="$OUT/app_$V"
CXX=/usr/bin/g++
case "$V" in k1) WV=WCFE_K1 ;; k31) WV=WCFE_K31 ;; kMD) WV=WCFE_KMD ;; *) echo "usage: $0 k1|k31|kMD"; exit 2 ;; esac
[ -f "$R/cellBP/main.cpp" ] || { echo "build_app.sh: $R not present here; using the prebuilt files if any"; exit 0; }
fresh=1
for dep in "$HERE/app_shim.cpp" "$HERE/build_app.sh" "$HERE/../wholecellfe_b200/csrc/ul_adapter.cpp" "$HERE/../include/wcfe.h"; do
    for out in "$OUT/cellBP_ref_$V" "$OUT/cellBP_gpu_$V"; do [ "$out" -nt "$dep" ] || fresh=0; done
done
[ "$fresh" = 1 ] && { echo "up to date: $OUT/cellBP_{ref,gpu}_$V"; exit 0; }
[ -f "$HERE/../wholecellfe_b200/libwcfe.so" ] || { echo "build_app.sh: build libwcfe.so first"; exit 1; }
mkdir -p "$OBJ"
FLAGS="-O3 -funroll-loops -w -D__LITTLE_ENDIAN__ -DUSE_FORTRAN_BLAS -DNO_GUI -fPIC"
INC="-I$R/cellBP -I$R/eltopo3d -I$R/common -I$R/common/newsparse -I$R/common/sparse -I$R/common/meshes -I$R/common/tunicate -I$R/cellBP/drivers"
# eltopo3d/Makefile.inc LIB_SRC + cellBP/Makefile SRC (UL_growth_explicit.cpp separately)
SRCS="eltopo3d/accelerationgrid.cpp eltopo3d/broadphasegrid.cpp eltopo3d/collisionpipeline.cpp eltopo3d/dynamicsurface.cpp
eltopo3d/edgecollapser.cpp eltopo3d/edgeflipper.cpp eltopo3d/edgesplitter.cpp eltopo3d/eltopo.cpp eltopo3d/impactzonesolver.cpp
eltopo3d/meshmerger.cpp eltopo3d/meshpincher.cpp eltopo3d/meshsmoother.cpp eltopo3d/meshrenderer.cpp eltopo3d/nondestructivetrimesh.cpp
eltopo3d/subdivisionscheme.cpp eltopo3d/surftrack.cpp eltopo3d/trianglequality.cpp
common/wallclocktime.cpp common/newsparse/sparse_matrix.cpp common/newsparse/krylov_solvers.cpp
common/tunicate/expansion.cpp common/tunicate/intersection.cpp common/tunicate/neg.cpp common/tunicate/orientation.cpp
common/root_parity_ccd_wrapper.cpp common/cubic_ccd_wrapper.cpp common/collisionqueries.cpp
cellBP/iomesh.cpp cellBP/lexer.cpp cellBP/newparser.cpp cellBP/geometryinit.cpp cellBP/scriptinit.cpp
common/bfstream.cpp common/makelevelset3.cpp common/marching_tiles_hires.cpp cellBP/main.cpp cellBP/UL_growth_explicit.cpp"
export CXX FLAGS INC R OBJ
echo $SRCS | tr ' ' '\n' | xargs -P "$(nproc)" -I{} sh -c 'o="$OBJ/$(basename {} .cpp).o"; [ "$o" -nt "$R/{}" ] || $CXX $FLAGS $INC -c "$R/{}" -o "$o"'
$CXX $FLAGS -I"$R/cellBP" -c "$HERE/app_shim.cpp" -o "$OBJ/app_shim.o"
$CXX $FLAGS $INC -I"$HERE/../include" -DWCFE_VARIANT=$WV -c "$HERE/../wholecellfe_b200/csrc/ul_adapter.cpp" -o "$OBJ/ul_adapter.o"
COMMON=$(ls "$OBJ"/*.o | grep -v -e UL_growth_explicit.o -e ul_adapter.o)
WRAPS="-Wl,--wrap=srand -Wl,--wrap=sprintf -Wl,--wrap=__sprintf_chk"
$CXX -o "$OUT/cellBP_ref_$V" $COMMON "$OBJ/UL_growth_explicit.o" $WRAPS -lstdc++ -lm
$CXX -o "$OUT/cellBP_gpu_$V" $COMMON "$OBJ/ul_adapter.o" $WRAPS -L"$HERE/../wholecellfe_b200" -lwcfe -Wl,-rpath,'$ORIGIN/../../wholecellfe_b200' -lstdc++ -lm
rm -rf "$OBJ"
echo "built $OUT/cellBP_ref_$V and $OUT/cellBP_gpu_$V"
