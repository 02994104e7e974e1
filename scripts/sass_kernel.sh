#!/bin/bash
# usage: scripts/sass_kernel.sh <mangled-substring> : rebuilds the cubin (ptxas -v) and dumps the flat SASS of the matching kernel to /tmp/an/kernel.sass
set -e
cd /root/repo/rustracer_b200/csrc
make ptxas 2>&1 | grep -A1 "Function properties for.*$1" | grep -E "Function|spill" | head -4
make ptxas 2>&1 | grep -B0 -A0 "Used.*registers" > /dev/null || true
FUN=$(cuobjdump -elf _build/rt_api.cubin | grep -o "_Z[A-Za-z0-9_]*$1[A-Za-z0-9_]*" | sort -u | head -1)
echo "kernel: $FUN"
mkdir -p /tmp/an
cuobjdump -sass -fun "$FUN" _build/rt_api.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^\s*//' | awk '{ $NF=""; print }' > /tmp/an/kernel.sass
wc -l /tmp/an/kernel.sass
awk '{op=$2; if (substr(op,1,1)=="@") op=$3; split(op,a,"."); c[a[1]]++} END{for(k in c) print c[k], k}' /tmp/an/kernel.sass | sort -rn | head -${2:-16} | tr '\n' ';'
echo
