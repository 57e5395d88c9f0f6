#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
run KV_TAG=gzero0
run KV_TAG=gzero25 KV_GZERO=0.25
run KV_TAG=gzero50 KV_GZERO=0.5
run KV_TAG=gzero50_noskip KV_GZERO=0.5 B200TRL_K1_SKIPZERO=0
