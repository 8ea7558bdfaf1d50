#!/bin/bash
# Records tests/golden/<prg>.ref.out: the reference's own example program, unchanged, linked with the compiled REFERENCE
# (oracle/_ref/prgs/<prg>_ref, built by scripts/build_prgs.sh) and run on the start files tests/common.py writes from the
# golden states.  Usage: bash tests/golden/make_prg_outputs.sh prg5 [prg2 ...]      (needs /root/reference at build time)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
for p in "$@"; do
  d=$(mktemp -d)
  (cd "$d" && python -c "import sys; sys.path[:0] = ['$ROOT', '$ROOT/tests']; import common as cm; cm.write_molecular_start_files('.')" \
      && "$ROOT/oracle/_ref/prgs/${p}_ref" > "$ROOT/tests/golden/$p.ref.out")
  rm -rf "$d"
  echo "recorded tests/golden/$p.ref.out"
done
