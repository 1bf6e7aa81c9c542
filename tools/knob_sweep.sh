#!/bin/bash
# usage: tools/knob_sweep.sh OUT "ENV1=a ENV2=b" "ENV1=c" ...   -- runs tools/prof_run.py 128 1 128 1 once per setting
out=$1; shift
: > $out
for kv in "$@"; do
  echo "== $kv" >> $out
  env $kv python tools/prof_run.py 128 1 128 1 2>&1 | tail -2 >> $out
done
