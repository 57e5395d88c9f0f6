#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
run KV_TAG=v151936
run KV_TAG=v50304 KV_V=50304 KV_B=32
run KV_TAG=v49152 KV_V=49152 KV_B=32
run KV_TAG=v100352 KV_V=100352
run KV_TAG=v128256 KV_V=128256
run KV_TAG=v524288 KV_V=524288 KV_B=4
run KV_TAG=v32000 KV_V=32000 KV_B=64
