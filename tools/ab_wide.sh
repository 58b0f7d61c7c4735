# A/B of the wide-kernel launch shape on one box: VARIANTS="name:flags ..." (flags separated by commas)
B=optical-rl-gym_b200/optical_rl_gym_b200/build.py
for v in $VARIANTS; do
  name=${v%%:*}; flags=${v#*:}
  ORLG_NVCC_EXTRA="$(echo $flags | tr ',' ' ')" python $B > /dev/null 2>&1
  for c in $CONFIGS; do
    echo "== $name $c"; python bench.py --config $c --steps 64 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
  done
done
