#!/bin/bash
# SASS of the dominant kernel of the shipped build (no GPU needed): full listing (hex encodings stripped) + mnemonic counts.
#   bash profiles/make_sass_listing.sh   ->  profiles/r2_tcf_sass.txt, profiles/r2_tcf_sass_mnemonics.txt
cd "$(dirname "$0")/.."
obj=pinn_elastodynamics_b200/csrc/pe_tcf.o
fun=$(cuobjdump -sass $obj | grep "Function :" | grep "resid_tcf_kernelILi5ELb0ELb0" | awk '{print $3}')
cuobjdump -sass -fun "$fun" $obj | sed -E 's#\s*/\* 0x[0-9a-f]+ \*/\s*$##' | grep -v '^\s*$' > profiles/r2_tcf_sass.txt
{
  echo "# resid_tcf_kernel<5, false, false> ($(grep -c '/\*[0-9a-f]\{4,\}\*/' profiles/r2_tcf_sass.txt) SASS instructions), mnemonic counts (tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, bulk TMA = UBLKCP, mbarrier = SYNCS, setmaxnreg = USETMAXREG)"
  grep -o '/\*[0-9a-f]\{4,\}\*/ *[@!UP0-9 ]*[A-Z][A-Z0-9_.]*' profiles/r2_tcf_sass.txt | sed -E 's#/\*[0-9a-f]+\*/ *(@!?U?P[0-9T]+ +)?##' | sed -E 's/\..*//' | sort | uniq -c | sort -rn
} > profiles/r2_tcf_sass_mnemonics.txt
head -30 profiles/r2_tcf_sass_mnemonics.txt
