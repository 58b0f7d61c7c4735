python -m pytest tests/test_rollout.py -x -q 2>&1 | tail -3
for T in 256 64 20; do python tools/time_rollout.py 65536 $T 10 2>&1 | tail -1; done
