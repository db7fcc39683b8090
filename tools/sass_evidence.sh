#!/bin/bash
# Regenerates profiles/<tag>_sass_evidence.txt and profiles/<tag>_ptxas.txt from the RELEASE build (run after build.py).
tag=${1:-r02}
so=neurofluid_b200/libnf_b200.so
sass=$(mktemp)
cuobjdump -sass $so > $sass
{
  echo "# SASS mnemonics proving the Blackwell-native paths (cuobjdump -sass $so), final build of the round"
  for m in UTCHMMA LDTM UBLKCP UTCBAR UTCATOMSWS "SYNCS.PHASECHK" "SYNCS.ARRIVE"; do echo "$m: $(grep -c "$m" $sass)"; done
  echo "HMMA (legacy mma.sync path, not counting UTCHMMA): $(grep -v UTCHMMA $sass | grep -c "HMMA")"
  echo
  echo "# kernels containing tensor-core MMA (UTCHMMA) / TMEM loads (LDTM) / bulk copies (UBLKCP): count, kernel"
  for m in UTCHMMA LDTM UBLKCP; do
    awk -v m=$m '/Function : /{f=$3} index($0, m){c[f]++} END{for (k in c) print m, c[k], k}' $sass | sort -k3
  done
} > profiles/${tag}_sass_evidence.txt
{
  echo "# ptxas -v summary of the final build of the round (registers / spills / smem per kernel)"
  for f in neurofluid_b200/build/*.ptxas.log; do
    awk '/Compiling entry function/{match($0, /'"'"'[^'"'"']+'"'"'/); fn=substr($0, RSTART+1, RLENGTH-2)} /bytes stack frame/{sp=$0; sub(/^ +/, "", sp)} /Used [0-9]+ registers/{u=$0; sub(/^ptxas info +: /, "", u); print fn ": " u " | " sp}' $f
  done
} > profiles/${tag}_ptxas.txt
rm -f $sass
wc -l profiles/${tag}_sass_evidence.txt profiles/${tag}_ptxas.txt
