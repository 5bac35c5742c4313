#!/bin/bash
# SASS evidence for profiles/: which Blackwell-specific instructions the built library contains.
#   bash tools/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
cd "$(dirname "$0")/.."
LIB=pano360_b200/libpano360_b200.so
cuobjdump -sass $LIB > /tmp/p360_sass.txt
echo "# cuobjdump -sass $LIB  ($(cuobjdump -lelf $LIB | head -n 3 | tr '\n' ' '))"
echo "# instruction counts over the whole library"
for I in UBLKCP SYNCS FFMA2 FMUL2 FADD2 UTMALDG LDGSTS "LDG.E.128" "STG.E.128" DFMA; do
  printf "%-10s %6d\n" "$I" "$(grep -c "$I" /tmp/p360_sass.txt)"
done
echo
echo "# per kernel: UBLKCP (TMA bulk copy global -> shared) / SYNCS (mbarrier) / FFMA2 (packed f32x2 FMA)"
awk '/Function :/ {name=$3} /UBLKCP/ {u[name]++} /SYNCS/ {s[name]++} /FFMA2/ {f[name]++}
     END {for (n in f) printf "%-110s UBLKCP %3d SYNCS %3d FFMA2 %4d\n", n, u[n], s[n], f[n]; for (n in u) if (!(n in f)) printf "%-110s UBLKCP %3d SYNCS %3d\n", n, u[n], s[n]}' /tmp/p360_sass.txt | sort
echo
echo "# blur_v_list_kernel: the bulk-copy staging loop (first UBLKCP and its surroundings)"
awk '/Function : _ZN4p36018blur_v_list_kernel/ {on=1} on && /UBLKCP/ && !done {for (i=NR-12;i<NR;i++) print buf[i%16]; print; tail=10; done=1; next} on && tail>0 {print; tail--} {buf[NR%16]=$0} /Function :/ && !/blur_v_list/ {on=0}' /tmp/p360_sass.txt | sed 's/^[ \t]*//' | cut -c1-140
