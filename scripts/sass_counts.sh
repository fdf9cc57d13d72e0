#!/usr/bin/env bash
# Evidence that the kernels are Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"): per compiled object, how often the
# SASS carries tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP), cp.async (LDGSTS) and the legacy mma.sync (HMMA).
#   scripts/sass_counts.sh > profiles/sass_counts.txt        (no GPU needed: cuobjdump reads the objects nvcc cross-compiled)
cd "$(dirname "$0")/.."
printf "%-16s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA UTCxMMA LDTM STTM UTMALDG UBLKCP LDGSTS HMMA
for o in build/obj/*.o; do
  n=$(basename "$o" | cut -d. -f1)
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-16s %8s %8s %8s %8s %8s %8s %8s %8s\n" "$n" "$(c UTCHMMA)" "$(c 'UTC[A-Z]*MMA')" "$(c LDTM)" "$(c STTM)" "$(c UTMALDG)" "$(c UBLKCP)" "$(c LDGSTS)" "$(c ' HMMA')"
done
