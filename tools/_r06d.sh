Q="--no-cpu-baseline --no-synthesize --no-vocoder --sustained-steps 60"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "stage_trace or golden or ten_step or weight_scale" > gpurun_out/r06d_tests.txt 2>&1; tail -3 gpurun_out/r06d_tests.txt
python tools/tail_timeline.py 64 344 2>&1 | tail -12
for i in 1 2; do
  timeout 300 python bench.py $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), round(d['config']['sustained']['value']/1e6,3), round(d['config']['serial']['value']/1e6,3), d['config5']['mel_sha256'][:8], round(d['config5']['valid_frames_per_s']/1e6,3))"
done
