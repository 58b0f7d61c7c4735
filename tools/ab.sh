B=optical-rl-gym_b200/optical_rl_gym_b200/build.py
python -m pytest tests/test_rollout.py -x -q 2>&1 | tail -3
python tools/time_rollout.py 65536 256 5 2>&1 | tail -1
python tools/time_rollout.py 65536 64 10 2>&1 | tail -1
python tools/time_rollout.py 65536 20 20 2>&1 | tail -1
for s in 24 64; do ORLG_RO_SPAN=$s python tools/time_rollout.py 65536 256 5 2>&1 | tail -1; done
ORLG_NVCC_EXTRA=-DORLG_PHASE_TIMING python $B > /dev/null 2>&1; python tools/rollout_phases.py 65536 256 2
