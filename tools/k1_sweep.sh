#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
for v in 151936 128256 65536 32000; do
  b=16; [ $v -lt 70000 ] && b=32; [ $v -lt 40000 ] && b=64
  for s in 9 5 4; do
    run KV_TAG=v${v}_bwdslots$s KV_V=$v KV_B=$b B200TRL_K1_BWD_SLOTS=$s
  done
done
