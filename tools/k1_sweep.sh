#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
for s in 11 10 9; do run KV_TAG=fused_slots$s B200TRL_K1_FUSED_SLOTS=$s; done
for s in 13 12 11; do run KV_TAG=wide_slots$s B200TRL_K1_GEOM=1 B200TRL_K1_FUSED_SLOTS=$s; done
