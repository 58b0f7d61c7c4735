#!/bin/bash
# instruction count of every kernel in liborlg.so whose name matches $1 (default: rollout)
LIB=${2:-optical-rl-gym_b200/optical_rl_gym_b200/liborlg.so}
cuobjdump -sass $LIB | awk -v pat="${1:-rollout}" '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) if (n ~ pat) print cnt[n], n}' | sort -n
