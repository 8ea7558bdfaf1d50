#!/bin/bash
# Compile the reference's own example programs, UNCHANGED, against include/sep.h + libsep.so.
# Sources are read where they lie (REF, default /root/reference); only binaries are written, into
# oracle/_ref/prgs/ (git-ignored test infrastructure, travels to the GPU box).  Also builds the same
# programs against the compiled reference (oracle/_ref/libsep_ref_fast.so) as *_ref for golden output.
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/oracle/_ref/prgs
mkdir -p "$OUT"
for p in prg0 prg1 prg2 prg3 prg4 prg5 prg6 prg7 prg8 prg9; do
  OMP=""; [ $p = prg5 ] && OMP="-fopenmp"      # prg5 calls the library from two OpenMP sections at once (prgs/prg5.c:56-70)
  gcc -std=c99 -O2 -w $OMP -I"$ROOT/include" "$REF/prgs/$p.c" -L"$ROOT/seplib_b200" -lsep -lm \
      -Wl,-rpath,'$ORIGIN/../../../seplib_b200' -o "$OUT/$p"
  echo "built $p against seplib-b200"
done
if [ -f "$ROOT/oracle/_ref/libsep_ref.so" ]; then
  for p in prg0 prg1 prg2 prg3 prg4 prg5 prg7 prg9; do
    gcc -std=c99 -O2 -w -DCOMPLEX -fopenmp -I"$REF/include" "$REF/prgs/$p.c" "$ROOT/oracle/_ref/libsep_ref.so" -lm \
        -Wl,-rpath,'$ORIGIN/..' -o "$OUT/${p}_ref"
    echo "built ${p}_ref against the reference"
  done
fi
