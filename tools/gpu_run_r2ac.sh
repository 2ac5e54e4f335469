mkdir -p gpurun_out
timeout 600 python bench.py --no-extra --no-learner > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ac_bench.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e']), json.dumps(d['stats_per_substep']))"
tail -3 gpurun_out/r2ac_bench.err
