set -x
python -m pytest tests/test_fused_gpu.py -m gpu -x -q -k "stream_of_device" 2>&1 | tail -15
python tools/time_stream.py T10 30 > gpurun_out/r02i_stream.txt 2>&1
cat gpurun_out/r02i_stream.txt
python bench.py --no-cpu > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -c 600 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_step_serial','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
PY
