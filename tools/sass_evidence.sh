#!/bin/bash
# SASS evidence (B200_PROFILING.md "What proves a Blackwell-native kernel"): per-kernel counts of the
# tcgen05 / TMA / TMEM mnemonics in libb200gemm.so.  usage: tools/sass_evidence.sh > profiles/rNN_sass_evidence.txt
cd "$(dirname "$0")/.."
SO="how-to-optimize-gemm_b200/libb200gemm.so"
echo "# cuobjdump -sass $SO  (sm_100a); counts per kernel of: UTC*MMA (tcgen05.mma), UTMALDG (TMA load),"
echo "# LDTM (tcgen05.ld), UTCBAR/SYNCS (mbarrier), FFMA2/FFMA, IMMA/HMMA (legacy mma.sync: must be 0)"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
/UTCHMMA/ {c["UTCHMMA"]++} /UTCQMMA/ {c["UTCQMMA"]++} /UTCIMMA/ {c["UTCIMMA"]++} /UTCOMMA/ {c["UTCOMMA"]++}
/UTMALDG/ {c["UTMALDG"]++} /UTMASTG/ {c["UTMASTG"]++} /LDTM/ {c["LDTM"]++} /UTCBAR/ {c["UTCBAR"]++}
/SYNCS\./ {c["SYNCS"]++} / FFMA2 / {c["FFMA2"]++} / FFMA / {c["FFMA"]++} / HMMA/ {c["HMMA"]++} / IMMA/ {c["IMMA"]++}
/UTCATOMSWS|UTCALLOC|UVIRTCOUNT/ {c["TMEM_ALLOC"]++}
function emit(   s,k) { s=""; n=split("UTCHMMA UTCQMMA UTCIMMA UTCOMMA UTMALDG LDTM UTCBAR SYNCS TMEM_ALLOC FFMA2 FFMA HMMA IMMA",ks," ");
  for (i=1;i<=n;i++) if (c[ks[i]]>0) s=s" "ks[i]"="c[ks[i]]; print name ":" s }
END { emit() }' | c++filt | sed 's/CUtensorMap_st, CUtensorMap_st, //' | cut -c1-230
echo
echo "# first tcgen05.mma site of the bf16 CTA-pair kernel (gemm_tc_kernel<0,256,6,float,ProdSingle,128,2>):"
cuobjdump -sass "$SO" | awk '/Function :.*gemm_tc_kernelILi0ELi256ELi6EfNS_10ProdSingleELi128ELi2E/{f=1} f&&/UTC.MMA/{print; n++} n>=4{exit}' | cut -c1-150
