#!/bin/bash
# Per-kernel census of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
# LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA, SYNCS = mbarrier, HMMA = mma.sync (warp-level tensor path), LDGSTS = cp.async.
SO=${1:-vrp-gym_b200/vrpx/libvrpx.so}
cuobjdump -sass "$SO" | c++filt | awk '
  /Function :/ { fn=$0; sub(/.*Function : /, "", fn); sub(/\(.*/, "", fn); if (!(fn in seen)) { seen[fn]=1; order[++n]=fn } }
  /UTCHMMA/ { a[fn,1]++ } /LDTM/ { a[fn,2]++ } /STTM/ { a[fn,3]++ } /UTMALDG/ { a[fn,4]++ } /SYNCS/ { a[fn,5]++ }
  /HMMA\.16816/ { a[fn,6]++ } /HMMA\.1688/ { a[fn,7]++ } /LDGSTS/ { a[fn,8]++ }
  END {
    printf "%-58s %8s %5s %5s %8s %6s %11s %10s %7s\n", "kernel", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "SYNCS", "HMMA.16816", "HMMA.1688", "LDGSTS"
    for (i = 1; i <= n; ++i) { f = order[i]; t = 0; for (j = 1; j <= 8; ++j) t += a[f,j]
      if (t) printf "%-58s %8d %5d %5d %8d %6d %11d %10d %7d\n", substr(f, 1, 58), a[f,1], a[f,2], a[f,3], a[f,4], a[f,5], a[f,6], a[f,7], a[f,8] } }'
