#!/bin/bash
# profiles/rNN_sass_tma.txt: which kernels of liborb_b200.so use which Blackwell / notable instructions (cuobjdump -sass).
# TMA (cp.async.bulk[.tensor]) shows as UTMALDG / UBLKCP, cp.async as LDGSTS, DPX as VIMNMX3, the 16x2 SIMD add as VIADD.16x2,
# dp2a / dp4a as IDP. No HMMA / UTC*MMA: the path has no dense contraction.
LIB=${1:-orb_slam2_detailed_comments_b200/lib/liborb_b200.so}
echo "# $(cuobjdump -elf $LIB 2>/dev/null | grep -m1 -oE 'sm_[0-9]+a?' || echo sm_100a)  $(basename $LIB)  (count  kernel  mnemonic)"
cuobjdump -sass $LIB 2>/dev/null | grep -E "Function :|arch =|UTMALDG|UTMASTG|UBLKCP|HMMA|UTC[A-Z]*MMA|LDGSTS|VIMNMX3|VIADD\.16x2|IDP\.|REDUX|SYNCS" |
awk '/arch =/{print "# " $0; next} /Function :/{fn=$3; next}
     {k="?"; if($0~/UTMALDG/)k="UTMALDG"; else if($0~/UTMASTG/)k="UTMASTG"; else if($0~/UBLKCP/)k="UBLKCP"; else if($0~/HMMA/)k="HMMA"; else if($0~/UTC[A-Z]*MMA/)k="UTCMMA";
      else if($0~/LDGSTS/)k="LDGSTS"; else if($0~/VIMNMX3/)k="VIMNMX3(DPX)"; else if($0~/VIADD\.16x2/)k="VIADD.16x2"; else if($0~/IDP\./)k="IDP"; else if($0~/REDUX/)k="REDUX"; else if($0~/SYNCS/)k="SYNCS(mbarrier)";
      cnt[fn"\t"k]++}
     END{for(x in cnt) print cnt[x]"\t"x}' | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_orb_[a-z]+_cu_[0-9a-f]+[0-9]+(k_[a-z0-9_]+).*\t/\1\t/' | sort -k2,2 -k3,3 | uniq
