#!/bin/bash
# Dump the SASS of the kernels in an object whose mangled name matches a pattern:  scripts/sass_dump.sh <obj> <pattern> <out>
OBJ=$1; PAT=$2; OUT=$3
FN=$(cuobjdump -elf $OBJ 2>/dev/null | grep -o "_ZN2mb[A-Za-z0-9_]*" | grep -E "$PAT" | grep -v _param_ | sort -u | head -1)
echo "function: $FN"
cuobjdump -sass -fun "$FN" $OBJ | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's|^\s+/\*([0-9a-f]{4})\*/\s+|\1 |; s|\s*/\* 0x[0-9a-f]+ \*/||; s|;\s*$||' > $OUT
wc -l $OUT
awk '{op=$2; if (op ~ /^@/) op=$3; split(op,a,"."); n[a[1]]++} END {for (k in n) print n[k], k}' $OUT | sort -rn | head -25 | tr '\n' ' '; echo
