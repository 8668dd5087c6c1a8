#!/bin/bash
# SASS evidence of the Blackwell-native paths: TMA bulk tensor loads (UTMALDG), mbarrier transactions (SYNCS),
# cluster barriers (UCGABAR), 128-bit shared / global accesses.  Writes profiles/r2_sass_evidence.txt.
set -e
cd "$(dirname "$0")/.."
LIB=evostencils_b200/csrc/libevostencils_b200.so
OUT=profiles/r2_sass_evidence.txt
SASS=$(mktemp)
cuobjdump -sass $LIB > $SASS
{
echo "# SASS evidence (cuobjdump -sass $LIB, built by __graft_entry__.build() with"
echo "# nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false); regenerate with scripts/sass_evidence.sh"
echo
echo "## instruction counts over the whole library"
grep -oE "UTMALDG\.[0-9]D|SYNCS\.[A-Z.0-9]+|LDS\.128|STG\.E\.128|LDG\.E\.128[.A-Z]*|BAR\.SYNC[.A-Z_]*|UCGABAR_[A-Z]+|ELECT|DFMA" $SASS | sort | uniq -c | sort -rn
echo
echo "## kernels with TMA loads: count of UTMALDG / mbarrier phase waits"
awk '/Function :/{fn=$3} /UTMALDG/{c[fn]++} /SYNCS.PHASECHK/{w[fn]++} END{for(f in c) print c[f], w[f], f}' $SASS | sort -k3 | while read a b f; do echo "$a UTMALDG  $b SYNCS.PHASECHK  $(echo $f | c++filt | sed 's/(.*//')"; done
echo
echo "## kernels with thread-block-cluster barriers"
awk '/Function :/{fn=$3} /UCGABAR_ARV/{c[fn]++} END{for(f in c) print c[f], f}' $SASS | sort -k2 | while read a f; do echo "$a UCGABAR_ARV  $(echo $f | c++filt | sed 's/(.*//')"; done
echo
echo "## one steady-state plane step of k3_rbgs_col<16, 8, true, 1> (default 513^3 smoother): barrier to barrier"
FN=$(grep "Function :" $SASS | awk '{print $3}' | grep "k3_rbgs_colILi16ELi8ELb1ELi1" | head -1)
awk -v fn="$FN" '/Function :/{on=($3==fn)} on' $SASS | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '/BAR.SYNC/{n++} n>=9 && n<10' | cut -c1-100
} > $OUT
rm -f $SASS
wc -l $OUT
