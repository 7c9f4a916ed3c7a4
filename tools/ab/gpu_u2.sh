#!/bin/bash
for v in main u2 main u2; do
  if [ $v = main ]; then L=""; else L="variants/lib_$v.so"; fi
  for m in trace score; do SARLACC_LIB=$L timeout 300 python tools/profile_forward.py 400000 a1 $m 4 | tail -1 | sed "s/^/$v /"; done
done
