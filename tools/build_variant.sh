#!/bin/bash
# Builds a variant of libsvgir_b200.so with extra nvcc flags into svg-ir_b200/svgir_b200/variants/<name>.so
# (kernel A/B runs: SVGIR_B200_LIB=<path> selects it, svgir_b200/_lib.py). usage: tools/build_variant.sh <name> "<flags>"
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/svgir_variant_$name
rm -rf $tmp; mkdir -p $tmp/svg-ir_b200 $root/svg-ir_b200/svgir_b200/variants
cp -r $root/svg-ir_b200/csrc $tmp/svg-ir_b200/csrc
cp -r $root/include $tmp/include
mkdir -p $tmp/svg-ir_b200/svgir_b200
rm -f $tmp/svg-ir_b200/csrc/*.o
make -C $tmp/svg-ir_b200/csrc -j8 EXTRA="$flags" > $tmp/build.log 2>&1 || { tail -20 $tmp/build.log; exit 1; }
cp $tmp/svg-ir_b200/svgir_b200/libsvgir_b200.so $root/svg-ir_b200/svgir_b200/variants/$name.so
echo built $root/svg-ir_b200/svgir_b200/variants/$name.so
