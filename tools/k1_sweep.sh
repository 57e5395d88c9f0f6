#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
for v in 151936 128256 65536 50304; do
  b=16; [ $v -lt 70000 ] && b=32
  run KV_TAG=v${v}_dense_dual0 KV_V=$v KV_B=$b B200TRL_K1_GEOM=3 B200TRL_K1_DUAL=0
  run KV_TAG=v${v}_dense_dual1 KV_V=$v KV_B=$b B200TRL_K1_GEOM=3 B200TRL_K1_DUAL=1
  run KV_TAG=v${v}_wide_dual0 KV_V=$v KV_B=$b B200TRL_K1_GEOM=1 B200TRL_K1_DUAL=0
  run KV_TAG=v${v}_wide_dual1 KV_V=$v KV_B=$b B200TRL_K1_GEOM=1 B200TRL_K1_DUAL=1
done
run KV_TAG=v32000_twin_dual0 KV_V=32000 KV_B=64 B200TRL_K1_DUAL=0
run KV_TAG=v32000_twin_dual1 KV_V=32000 KV_B=64 B200TRL_K1_DUAL=1
run KV_TAG=v32000_dense_dual1 KV_V=32000 KV_B=64 B200TRL_K1_DUAL=1 B200TRL_K1_GEOM=3
