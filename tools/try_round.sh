#!/bin/bash
# GPU gate for a change: the whole -m gpu suite, the B=256 launch table and bench.py under environment variants
# tools/try_round.sh <tag> "VAR=1" "VAR2=1 VAR3=x" ...   (writes gpurun_out/<tag>_*)
tag=$1; shift
o=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $o/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $o/${tag}_pytest_gpu.log
tail -4 $o/${tag}_pytest_gpu.log
grep -q "rc=0" $o/${tag}_pytest_gpu.log || exit 1
timeout 300 python tools/profile_solve.py 256 344 > $o/${tag}_launch_table_B256_T344.txt 2>&1
grep "s0\.\|s1\.\|final\|total" $o/${tag}_launch_table_B256_T344.txt
timeout 1500 tools/bench_variants.sh "$@" > $o/${tag}_variants.txt 2>&1
cat $o/${tag}_variants.txt
