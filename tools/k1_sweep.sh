#!/bin/bash
# A/B of the K1 resident kernel's launch knobs (one JSON line per variant; see tools/k1_variants.py for the fields).
#   gpurun -- 'bash tools/k1_sweep.sh > gpurun_out/k1_sweep.jsonl'
# The policy in pick_geom() (csrc/k1_resident.cu) and the numbers quoted in DESIGN.md section 3 come from runs of this
# script with the blocks below edited in and out; the default set re-measures the decisions that matter most.
run() { env "$@" python tools/k1_variants.py; }

# 1. fused pass: consumer count / chunk size / ring depth per vocabulary (1 wide 512, 3 dense 768, 4 mid 640, 2 twin)
for v in 151936 128256 100352 65536 50304 49152 32000; do
  b=16; [ $v -lt 70000 ] && b=32; [ $v -lt 40000 ] && b=64
  run KV_TAG=v${v}_auto KV_V=$v KV_B=$b
  for gm in 1 3 4; do run KV_TAG=v${v}_geom$gm KV_V=$v KV_B=$b B200TRL_K1_GEOM=$gm; done
done
# 2. backward-only ring depth (short rings keep an SM's read and write streams together)
for s in 9 5 4; do run KV_TAG=bwd_slots$s B200TRL_K1_BWD_SLOTS=$s; done
# 3. streaming vs clustered forward-only / backward-only, accumulation chains, dlogits via TMA stores vs registers
run KV_TAG=fwd_cluster2 B200TRL_K1_FWD_CS=2 B200TRL_K1_GEOM=3
run KV_TAG=bwd_cluster2 B200TRL_K1_BWD_CS=2
run KV_TAG=fused_bulk_stores B200TRL_K1_DIRECT=0
# 4. large vocabularies: 4- and 8-CTA clusters against the row kernel
for v in 200000 262144 524288; do
  run KV_TAG=v${v}_resident KV_V=$v KV_B=8
  run KV_TAG=v${v}_row KV_V=$v KV_B=8 KV_ROW=1
done
