Q="--no-cpu-baseline --no-synthesize --no-vocoder --no-config5 --sustained-steps 60"
for v in "MTTS_X=0 4" "MTTS_X=0 3" "MTTS_X=0 5" "MTTS_X=0 6" "MTTS_PAIR_MIN_CHUNKS=4 4" "MTTS_PAIR_MIN_CHUNKS=12 4" "MTTS_GNBQKV=1 4" "MTTS_QKV_GEMM=1 4" "MTTS_X=0 4"; do
  set -- $v
  env $1 timeout 300 python bench.py $Q --in-flight $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), round(d['config']['sustained']['value']/1e6,3), round(d['config']['serial']['value']/1e6,3))"
done > gpurun_out/r05s_ab.txt 2>&1
cat gpurun_out/r05s_ab.txt
for l in 4 5 6; do python tools/config5.py 4096 22016 $l 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('config5 lanes $l', round(d['valid_frames_per_s']/1e6,3), d['mel_sha256'][:8])"; done
