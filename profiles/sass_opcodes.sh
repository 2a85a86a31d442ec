#!/bin/bash
# Per-kernel SASS opcode counts of the built library: the evidence that the contraction kernels are tcgen05 / TMEM / TMA
# (UTC*MMA, LDTM / STTM, UTMALDG / UTMASTG) and that none falls back to the legacy HMMA path.
#   bash profiles/sass_opcodes.sh > profiles/r2_sass_opcodes.txt
SO=${1:-auto_differentiation_b200/libb200ad.so}
/usr/local/cuda/bin/cuobjdump -sass "$SO" | awk '
  /Function : / { fn=$3 }
  { for (i=1;i<=NF;i++) { t=$i; sub(/[;,].*/,"",t);
      if (t ~ /^UTC[A-Z]*MMA/) m[fn]++;
      else if (t ~ /^LDTM/) l[fn]++;
      else if (t ~ /^STTM/) s[fn]++;
      else if (t ~ /^UTMALDG/) tl[fn]++;
      else if (t ~ /^UTMASTG/) ts[fn]++;
      else if (t ~ /^UBLKCP/) bc[fn]++;
      else if (t ~ /^HMMA/) h[fn]++;
      else if (t == "BRA.U.ANY") bu[fn]++;
      else if (t ~ /^SYNCS/) sy[fn]++;
      else if (t ~ /^UCGABAR/) cg[fn]++;
  } }
  /Function : / { seen[fn]=1 }
  END { printf "%-8s %-6s %-6s %-8s %-8s %-7s %-5s %-9s %-6s %-7s %s\n","UTC*MMA","LDTM","STTM","UTMALDG","UTMASTG","UBLKCP","HMMA","BRA.U.ANY","SYNCS","UCGABAR","kernel";
        for (f in seen) printf "%-8d %-6d %-6d %-8d %-8d %-7d %-5d %-9d %-6d %-7d %s\n", m[f],l[f],s[f],tl[f],ts[f],bc[f],h[f],bu[f],sy[f],cg[f],f }' | (read -r hdr; echo "$hdr"; sort -k1,1nr -k11)
