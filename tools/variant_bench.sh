#!/bin/bash
# A/B of library variants (tools/variants/*.so) on a 256 MiB shard: prints value and the kernel time
B="python bench.py --size 268435456 --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 8192 --compress-sample 1048576"
echo "default: $($B 2>/dev/null | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print(round(l["value"],1), round(l["roofline"]["kernel_ms"],1))')"
for v in "$@"; do
  echo "$v: $(SQZ_B200_LIB=$PWD/tools/variants/$v $B 2>/dev/null | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print(round(l["value"],1), round(l["roofline"]["kernel_ms"],1))')"
done
