# A/B build of the library with ONE source file replaced or recompiled with extra flags, for same-box comparisons
# (tools/ab_lib.py runs any script of this repo against it):
#   tools/build_ab.sh <name> <file.cu in csrc, or a path to an alternate version of it> [extra nvcc flags]  ->  tools/ab/libe2b_<name>.so
# e.g. the previous commit's kernel:  git show HEAD~1:video-to-audio-and-piano-rp_b200/csrc/elementwise.cu > /tmp/elementwise.cu
#                                     tools/build_ab.sh old /tmp/elementwise.cu
set -e
cd "$(dirname "$0")/.."
PK=video-to-audio-and-piano-rp_b200
name=$1; src=$2; shift 2
base=$(basename "$src" .cu)
[ -f "$src" ] || src=$PK/csrc/$base.cu
python $PK/build.py > /dev/null
mkdir -p tools/ab
cp "$src" $PK/csrc/_ab_$base.cu                      # compiled from csrc/ so that its includes resolve
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $PK/csrc/_ab_$base.cu -o tools/ab/${base}_$name.o
rm $PK/csrc/_ab_$base.cu
nvcc -shared -o tools/ab/libe2b_$name.so $(ls $PK/build/*.o | grep -v "/$base.o") tools/ab/${base}_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static -ldl
ls -la tools/ab/libe2b_$name.so
