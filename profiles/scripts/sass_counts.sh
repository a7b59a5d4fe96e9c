#!/bin/bash
# SASS evidence per kernel of libhydra_pspec_b200.so: FP64 tensor instructions (DMMA.8x8x4), TMA bulk
# copies (UBLKCP), mbarrier operations (SYNCS), cp.async (LDGSTS), and the absence of tcgen05 (UTC*MMA) -- tcgen05 has no FP64
# kind, DMMA is Blackwell's FP64 tensor instruction.
#   bash profiles/scripts/sass_counts.sh > profiles/sass_counts.txt
cd "$(dirname "$0")/../.."
SO=hydra_pspec_b200/csrc/libhydra_pspec_b200.so
echo "# cuobjdump -sass $SO ($(date -u +%Y-%m-%d)), instruction counts per kernel"
printf "%-60s %7s %7s %7s %7s %7s %7s %7s\n" kernel DMMA UBLKCP SYNCS LDGSTS LDG LDS UTCxMMA
cuobjdump -sass $SO | awk '
/Function : /{ if (name != "") printf "%-60s %7d %7d %7d %7d %7d %7d %7d\n", name, d, u, s, l, g, h, t; name=$3; d=u=s=l=g=h=t=0 }
/DMMA/{d++} /UBLKCP/{u++} /SYNCS/{s++} /LDGSTS/{l++} / LDG\./{g++} / LDS/{h++} /UTC[A-Z]*MMA/{t++}
END{ printf "%-60s %7d %7d %7d %7d %7d %7d %7d\n", name, d, u, s, l, g, h, t }' | sed 's/_ZN2hp//' | c++filt 2>/dev/null | sort
