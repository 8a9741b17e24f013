#!/bin/bash
# Per-kernel SASS instruction counts of the shipping library: the Blackwell-native tell (B200_PROFILING.md):
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKRED/UBLKCP = bulk
#   reduce/copy, HMMA = legacy mma.sync (must be 0).  usage: scripts/sass_summary.sh [libfa_b200.so] > profiles/rNN_sass_counts.txt
LIB=${1:-$(dirname "$0")/../flash-attention-turing_b200/flash_attn_turing/libfa_b200.so}
echo "# SASS mnemonic counts per kernel of $(basename $LIB) ($(date -u +%Y-%m-%d)); cuobjdump -sass"
printf "%-96s %8s %6s %6s %8s %8s %8s %6s %6s %8s\n" kernel UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKRED REDG HMMA MUFU.EX2
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") flush(); name=$3; split("",c) ; next }
  { if ($0 ~ /UTCHMMA/) c["mma"]++; if ($0 ~ /LDTM/) c["ldtm"]++; if ($0 ~ /STTM/) c["sttm"]++;
    if ($0 ~ /UTMALDG/) c["ldg"]++; if ($0 ~ /UTMASTG/) c["stg"]++; if ($0 ~ /UBLKRED/) c["red"]++;
    if ($0 ~ /[^U]REDG|RED\.E/) c["redg"]++; if ($0 ~ / HMMA/) c["hmma"]++; if ($0 ~ /MUFU\.EX2/) c["ex2"]++ }
  function flush() { printf "%s %8d %6d %6d %8d %8d %8d %6d %6d %8d\n", name, c["mma"], c["ldtm"], c["sttm"], c["ldg"], c["stg"], c["red"], c["redg"], c["hmma"], c["ex2"] }
  END { if (name != "") flush() }' | while read -r line; do
    n=$(echo "$line" | awk '{print $1}'); d=$(echo "$n" | c++filt | sed 's/(CUtensorMap_st.*//; s/(fa100::BwdParams.*//; s/^void //; s/(bool)//g; s/(int)//g'); echo "$line" | awk -v d="$d" '{ $1=""; printf "%-96s%s\n", substr(d,1,95), $0 }'
  done | sort
