#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05.mma (UTCHMMA),
# tcgen05.ld (LDTM), tcgen05.commit (UTCBAR), TMA (UTMALDG), cp.async (LDGSTS), mbarrier (SYNCS), setmaxnreg
# (USETMAXREG), packed fp32 FMA (FFMA2).   usage: scripts/sass_counts.sh > profiles/r02_sass_counts.txt
cd "$(dirname "$0")/.."
so=hybrid-als-twotower-recommender_b200/libhals_b200.so
echo "# $(git rev-parse --short HEAD) $(date -u +%F) cuobjdump -sass $so"
cuobjdump -sass $so | awk '
/Function :/ { f=$3 }
/UTCHMMA/ {a[f]++; k[f]=1} /LDTM/ {b[f]++; k[f]=1} /UTCBAR/ {c[f]++; k[f]=1} /UTMALDG/ {d[f]++; k[f]=1}
/LDGSTS/ {e[f]++; k[f]=1} /SYNCS/ {g[f]++; k[f]=1} /USETMAXREG/ {h[f]++; k[f]=1} /FFMA2/ {m[f]++; k[f]=1}
END { printf "%-8s %-6s %-7s %-8s %-7s %-6s %-10s %-6s %s\n","UTCHMMA","LDTM","UTCBAR","UTMALDG","LDGSTS","SYNCS","USETMAXREG","FFMA2","kernel";
      for (f in k) printf "%-8d %-6d %-7d %-8d %-7d %-6d %-10d %-6d %s\n", a[f],b[f],c[f],d[f],e[f],g[f],h[f],m[f],f }' | (read -r hdr; echo "$hdr"; sort -k9) | while read -r line; do
  set -- $line; name=$(echo "$9" | c++filt 2>/dev/null | cut -c1-110); printf "%-8s %-6s %-7s %-8s %-7s %-6s %-10s %-6s %s\n" "$1" "$2" "$3" "$4" "$5" "$6" "$7" "$8" "$name"; done
