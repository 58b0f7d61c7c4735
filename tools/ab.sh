B=optical-rl-gym_b200/optical_rl_gym_b200/build.py
C=optical-rl-gym_b200/csrc
for v in $VARIANTS; do
  cp exp/$v/orlg_rollout.cuh exp/$v/orlg_api.cu $C/
  python $B > /dev/null 2>&1
  echo "== $v"; for T in 256 20; do python tools/time_rollout.py 65536 $T 10 2>&1 | tail -1; done
done
