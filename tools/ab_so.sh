#!/bin/bash
# A/B of prebuilt library variants on one GPU box: VARIANTS="v0 v1" [TESTS=1] [ALLKINDS=1] tools/ab_so.sh
L=optical-rl-gym_b200/optical_rl_gym_b200/liborlg.so
for v in $VARIANTS; do
  cp exp/$v/liborlg.so $L; touch $L
  echo "== $v"
  for T in 256 20; do python tools/time_rollout.py 65536 $T 10 2>&1 | tail -1; done
  if [ -n "$ALLKINDS" ]; then
    KIND=RMSA-v0 POLICY=sap_ff python tools/time_rollout.py 65536 64 7 2>&1 | tail -1
    KIND=RWA-v0 POLICY=sap_ff python tools/time_rollout.py 65536 64 7 2>&1 | tail -1
  fi
  if [ -n "$TESTS" ]; then python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -3; fi
done
