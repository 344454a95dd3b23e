#!/bin/bash
# builds libprt_b200 variants with different -D knobs into build_variants/<name>.so  (usage: name "flags" ...)
set -e
cd "$(dirname "$0")/../physics-based-ray-tracing_b200/csrc"
mkdir -p ../../build_variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  rm -f *.o
  make -s -j8 EXTRA="$flags" OUT=../../build_variants/$name.so >/dev/null
  echo "built $name ($flags)"
done
rm -f *.o
make -s -j8 >/dev/null
