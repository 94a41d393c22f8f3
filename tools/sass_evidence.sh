#!/bin/bash
# Which kernels of liblz77b200.so use the TMA bulk copy / mbarrier / warp-match / warp-reduce
# instructions the design claims (SASS mnemonics, sm_100a):
#   UBLKCP.S.G  cp.async.bulk global -> shared (TMA load)    UBLKCP.G.S  shared -> global (TMA store)
#   SYNCS       mbarrier arrive / try_wait                    MATCH.ANY   __match_any_sync
#   REDUX       __reduce_{max,min}_sync                       VOTE        __ballot/__any_sync
#   ATOMS       shared-memory atomics                         STG.E.128   128-bit global stores
# usage: tools/sass_evidence.sh [lib]  > profiles/rNN_sass_evidence.txt
LIB=${1:-lz77_b200/liblz77b200.so}
echo "# cuobjdump -sass $LIB (sm_100a): instruction counts per kernel"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { name=$3; sub(/^_ZN4lz77[0-9]*/, "", name); names[++n]=name; next }
  name != "" {
    if ($0 ~ /UBLKCP\.S\.G/) c[name,"UBLKCP.S.G"]++
    if ($0 ~ /UBLKCP\.G\.S/) c[name,"UBLKCP.G.S"]++
    if ($0 ~ /SYNCS/) c[name,"SYNCS"]++
    if ($0 ~ /MATCH\.ANY/) c[name,"MATCH.ANY"]++
    if ($0 ~ /REDUX/) c[name,"REDUX"]++
    if ($0 ~ /VOTE/) c[name,"VOTE"]++
    if ($0 ~ /ATOMS/) c[name,"ATOMS"]++
    if ($0 ~ /STG\.E\.128/) c[name,"STG.E.128"]++
    if ($0 ~ /LDS\.128/) c[name,"LDS.128"]++
    if ($0 ~ /HMMA|UTCHMMA|UTCMMA|IMMA/) c[name,"tensor"]++
  }
  END {
    split("UBLKCP.S.G UBLKCP.G.S SYNCS MATCH.ANY REDUX VOTE ATOMS STG.E.128 LDS.128 tensor", cols, " ")
    printf "%-64s", "kernel"; for (j=1;j<=10;j++) printf " %10s", cols[j]; printf "\n"
    for (i=1;i<=n;i++) { printf "%-64s", substr(names[i],1,64); for (j=1;j<=10;j++) printf " %10d", c[names[i],cols[j]]+0; printf "\n" }
  }'
