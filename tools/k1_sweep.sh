#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
for v in 151936 152064 128256 100352 65536; do
  b=16; [ $v -lt 70000 ] && b=32
  run KV_TAG=v${v}_wide512 KV_V=$v KV_B=$b B200TRL_K1_GEOM=1
  run KV_TAG=v${v}_576 KV_V=$v KV_B=$b B200TRL_K1_GEOM=5
  run KV_TAG=v${v}_640 KV_V=$v KV_B=$b B200TRL_K1_GEOM=4
  run KV_TAG=v${v}_704 KV_V=$v KV_B=$b B200TRL_K1_GEOM=6
  run KV_TAG=v${v}_768 KV_V=$v KV_B=$b B200TRL_K1_GEOM=3
done
