#!/bin/bash
# tools/build_variant.sh NAME "extra nvcc flags": an A/B build of libdre.so into variants/libdre_NAME.so (DRE_LIB selects it)
set -e
NAME=$1; FLAGS=$2
D=/tmp/dre_variant_$NAME; rm -rf $D; mkdir -p $D
cp -r dr-mpc_b200 include $D/
sed -i "s|COMMON = \[|COMMON = [$(for f in $FLAGS; do printf '"%s", ' $f; done)|" $D/dr-mpc_b200/build.py
(cd $D && python dr-mpc_b200/build.py --force > /dev/null)
mkdir -p variants; cp $D/dr-mpc_b200/libdre.so variants/libdre_$NAME.so; echo built variants/libdre_$NAME.so
