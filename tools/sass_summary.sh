#!/bin/sh
# SASS mnemonic counts per object (tensor-core, TMA, barrier and FP64 instructions): profiles/r02_sass_summary.txt.
# Needs only cuobjdump (no GPU).   sh tools/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/../pywgcna_b200/csrc/build" || exit 1
for o in tom_i8.o corr_i8.o corr.o tom.o sft.o linkage.o; do
  echo "== $o"
  cuobjdump -sass "$o" | grep -oE '\b(UTC[A-Z0-9_.]+|UTMA[A-Z0-9_.]+|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|DMMA[A-Z0-9_.]*|DFMA[A-Z.]*|DADD|DMUL[A-Z.]*|SYNCS[A-Z0-9_.]*|MEMBAR[A-Z.]*|CCTL[A-Z.]*|UCGABAR_[A-Z]+)\b' \
    | sort | uniq -c | sort -rn
done
