#!/bin/bash
for t in "" 256 512 1024; do
  if [ -z "$t" ]; then python tools/single_case.py; else HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_cs$t.so python tools/single_case.py; fi
done
