#!/bin/bash
# Round artefacts on one B200 (run under gpurun; writes gpurun_out/<tag>_*): tools/collect_profiles.sh <tag>
tag=${1:-r06z}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/${tag}_gpu_tests.txt 2>&1; tail -2 $o/${tag}_gpu_tests.txt
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
X="--no-cpu-baseline --no-synthesize --no-vocoder --no-config5 --sustained-steps 0 --steps 10"
for n in 2 4 50; do python bench.py --n-timesteps $n $X > $o/${tag}_bench_n$n.json 2>> $o/${tag}_bench.err; done
python bench.py --ragged $X > $o/${tag}_bench_ragged.json 2>> $o/${tag}_bench.err
python bench.py --batch 16 --frames 2048 $X > $o/${tag}_bench_B16_T2048.json 2>> $o/${tag}_bench.err
python bench.py --batch 16 --frames 2048 --ragged $X > $o/${tag}_bench_B16_T2048_ragged.json 2>> $o/${tag}_bench.err
python bench.py --batch 1 --frames 344 $X > $o/${tag}_bench_B1.json 2>> $o/${tag}_bench.err
python tools/profile_solve.py 256 344 > $o/${tag}_launch_table_B256_T344.txt 2>&1
python tools/sweep_nsub.py 64 344 10 1,2 > $o/${tag}_chain_sweep.txt 2>&1
python tools/sweep_nsub.py 192 344 10 1 >> $o/${tag}_chain_sweep.txt 2>&1
# ncu: launch list of one eager solve (cold, serialised), then a full capture of stage 0 + 1 of the second evaluation
python tools/one_solve.py 64 344 2 > $o/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|gn_apply|attention|ff_tail|qkv" -c 120 --csv --log-file $o/${tag}_launches.csv python tools/one_solve.py 64 344 2 > $o/${tag}_ncu1.log 2>&1
python tools/one_solve.py 256 344 1 > $o/${tag}_plain256.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|gn_apply|attention|ff_tail|qkv" -s 49 -c 17 -o $o/${tag}_stage01 python tools/one_solve.py 256 344 1 > $o/${tag}_ncu2.log 2>&1
ls -la $o | grep ${tag}
