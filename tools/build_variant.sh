# tools/build_variant.sh <name> <extra nvcc flags...>: liblsim.so built with extra -D flags into build_variants/<name>.so
# (kernel experiments; tools/gpu_variants_r2.sh runs them against the shipped library on one box via LSIM_LIB)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build_variants/$NAME
mkdir -p $OUT
cd $ROOT/jaxtronomy_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -static-global-template-stub=false -Xcompiler -fPIC -Xcompiler -O2 $*"
nvcc $FLAGS -c -o $OUT/lsim_hot.o lsim_hot.cu &
nvcc $FLAGS --split-compile 8 -c -o $OUT/lsim_api.o lsim_api.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o $ROOT/build_variants/$NAME.so $OUT/lsim_hot.o $OUT/lsim_api.o
echo built $ROOT/build_variants/$NAME.so
