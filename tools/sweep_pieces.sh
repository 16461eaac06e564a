#!/bin/bash
# Pieces of a large shard (launch_v2) and CTAs per SM of the overlapped phase 2: -DSQZ_TUNING build
B="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --cpu-sample 8192 --compress-sample 1048576 --kind-bytes 0"
run() {
  echo "size $1 pieces $2 fin_ctas $3: $(SQZ_B200_LIB=$PWD/tools/variants/tuning.so SQZ_PIECES=$2 SQZ_FIN_CTAS=$3 $B --size $1 2>/dev/null | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print(round(l["value"],1), round(l["roofline"]["kernel_ms"],1), round(l["e2e"]["value"],1), l["parity_check"]["table_equals_oracle_B"])')"
}
for cfg in "1 2" "8 2" "16 1" "16 2" "16 3" "32 2"; do run 1073741824 $cfg; done
for cfg in "1 2" "2 2" "4 2" "8 2"; do run 268435456 $cfg; done
