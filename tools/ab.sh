ORLG_NVCC_EXTRA=-DORLG_PHASE_TIMING python optical-rl-gym_b200/optical_rl_gym_b200/build.py > /dev/null 2>&1; python tools/rollout_phases.py 65536 256 2;  python tools/rollout_phases.py 65536 20 10
