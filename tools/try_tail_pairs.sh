#!/bin/bash
# parity + timing of the CTA-pair tail kernel (MTTS_TAIL_PAIRS=1) next to the default; writes gpurun_out/<tag>_*
tag=${1:-r01i}
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "opt_in_variants and PAIRS" > $o/${tag}_tailpair_pytest.log 2>&1
echo "pytest rc=$?" >> $o/${tag}_tailpair_pytest.log
tail -5 $o/${tag}_tailpair_pytest.log
grep -q "rc=0" $o/${tag}_tailpair_pytest.log || exit 1
timeout 300 python tools/profile_solve.py 256 344 > $o/${tag}_launch_table_B256_T344.txt 2>&1
MTTS_TAIL_PAIRS=1 timeout 300 python tools/profile_solve.py 256 344 > $o/${tag}_launch_table_B256_T344_tailpairs.txt 2>&1
MTTS_TAIL_PAIRS=1 timeout 300 python tools/tail_timeline.py > $o/${tag}_tail_timeline_pairs.txt 2>&1
timeout 1200 tools/bench_variants.sh "MTTS_X=0" "MTTS_TAIL_PAIRS=1" "MTTS_TAIL_PAIRS=1 MTTS_PAIRS=1" > $o/${tag}_variants.txt 2>&1
cat $o/${tag}_variants.txt
grep -i "tail" $o/${tag}_launch_table_B256_T344.txt | head -8
grep -i "tail" $o/${tag}_launch_table_B256_T344_tailpairs.txt | head -8
