#!/bin/bash
# Per-kernel SASS mnemonic counts of the built library (evidence for the instruction mix: bulk async copies UBLKCP,
# mbarrier SYNCS, FP64 tensor-core DMMA, DFMA ...).  usage: tools/sass_listing.sh > profiles/sass_rNN.md
so=pino_locoman_b200/libpinolocoman_b200.so
echo "# SASS mnemonic counts per kernel ($(basename $so), $(date -u +%Y-%m-%d), cuobjdump -sass)"
echo
echo "| kernel | instructions | DFMA | DMUL | DADD | DMMA | UBLKCP | UBLKPF | SYNCS | BAR | LDS | STS | LDG | STG | SHFL | ATOMS | MUFU |"
echo "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"
cuobjdump -sass $so | awk '
/Function :/ { if (name != "") flush(); name=$3; n=0; delete c; next }
/^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/ {
  op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); c[op]++; n++
}
function flush() {
  dn=""; cmd="c++filt " name; cmd | getline dn; close(cmd); sub(/\(.*/, "", dn); if (dn == "") dn=name;
  printf("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", dn, n, c["DFMA"], c["DMUL"], c["DADD"], c["DMMA"], c["UBLKCP"], c["UBLKPF"], c["SYNCS"], c["BAR"], c["LDS"], c["STS"], c["LDG"], c["STG"], c["SHFL"], c["ATOMS"], c["MUFU"])
}
END { if (name != "") flush() }'
