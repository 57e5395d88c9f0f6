#!/bin/bash
run() { env "$@" python tools/k1_variants.py; }
run KV_TAG=default
run KV_TAG=poly1 B200TRL_K1_POLY=1
run KV_TAG=poly2 B200TRL_K1_POLY=2
run KV_TAG=default_again
