#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
run KV_TAG=default
run KV_TAG=skipzero_off B200TRL_K1_SKIPZERO=0
