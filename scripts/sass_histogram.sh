#!/bin/bash
# Opcode histogram per kernel of libtbn_sm100.so (cuobjdump -sass): the Blackwell-native evidence (UTCHMMA = tcgen05.mma,
# LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, UTCCP = tcgen05.cp).
LIB=attention_based_tbn_b200/lib/libtbn_sm100.so
OUT=${1:-profiles/sass_opcodes.txt}
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep -o 'V[0-9][0-9.]*' | tail -1))"
  echo "# whole library: tensor-core / TMA / TMEM opcodes"
  cuobjdump -sass $LIB | grep -oE "\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|UTCATOMSWS|HMMA|IMMA|ELECT|UCGABAR_ARV|UCGABAR_WAIT)(\.[A-Z0-9_.]+)?" | sed 's/\..*//' | sort | uniq -c | sort -rn
  echo
  echo "# per kernel (function : opcode counts)"
  cuobjdump -sass $LIB | awk '
    /Function : / { fn=$3 }
    { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|HMMA|FFMA|LDG|STG|LDS|STS)(\.|;|$)/) { op=$i; sub(/[.;].*/,"",op); c[fn" "op]++ } }
    END { for (k in c) print c[k], k }' | sort -k2,2 -k1,1rn | awk '{ if ($2!=last) { if (last!="") print line; line=$2" :"; last=$2 } line=line" "$3"="$1 } END { print line }' | c++filt | cut -c1-260
} > $OUT
wc -l $OUT
