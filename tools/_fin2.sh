timeout 60 compute-sanitizer --tool racecheck python tools/png_sanitize.py 2>&1 | tail -2
timeout 60 compute-sanitizer --tool memcheck python tools/png_sanitize.py 2>&1 | tail -2
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_refimg.json 2> gpurun_out/bench_refimg.err
python -c "
import json; d=json.load(open('gpurun_out/bench_refimg.json')); print(d['value'], d['parity'].get('reference_image'))"
