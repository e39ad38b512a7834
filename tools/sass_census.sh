#!/bin/bash
# Instruction census of the product library (the evidence that the hot ops are tcgen05 / TMEM / TMA code):
#   bash tools/sass_census.sh > profiles/rNN_sass_census.txt
LIB=${1:-asr-2pass_b200/lib/libb200pf.so}
echo "# cuobjdump -sass $LIB | mnemonic counts ($(date -u +%Y-%m-%d))"
cuobjdump -sass "$LIB" > /tmp/sass_census.$$ || exit 1
for m in 'UTCHMMA\.2CTA' 'UTCHMMA ' 'UTCBAR\.2CTA\.MULTICAST' 'UTCBAR ' 'LDTM' 'STTM' 'UTMALDG\.2D\.2CTA' 'UTMALDG\.2D ' 'UTMASTG\.2D' \
         'UTMAREDG\.2D\.ADD' 'UTCATOMSWS' 'UCGABAR_ARV' 'UCGABAR_WAIT' 'ELECT' 'BRA\.U\.ANY' 'SYNCS\.PHASECHK' 'SYNCS\.ARRIVE' 'HMMA\.16816' \
         'HMMA\.1688' 'LDSM' 'FFMA2' 'MUFU\.EX2' 'REDG?\.E\.ADD' 'LDGSTS' 'DFMA'; do
  printf "%-28s %6d\n" "$(echo "$m" | sed 's/\\//g')" "$(grep -cE "$m" /tmp/sass_census.$$)"
done
echo "# kernels"
grep "Function :" /tmp/sass_census.$$ | grep -oE "[0-9]+[a-z][a-z_0-9]*kernel" | sed -E 's/.*[0-9]+([a-z][a-z_]*[a-z0-9_]*kernel)$/\1/' | sort | uniq -c | sort -rn
rm -f /tmp/sass_census.$$
