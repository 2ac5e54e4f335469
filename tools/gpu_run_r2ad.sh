mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity.py -m gpu -q -k "expert_action" 2>&1 | tail -2
timeout 600 python bench.py --no-extra --no-learner > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ad_bench.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e']), json.dumps(d['stats_per_substep']))"
tail -3 gpurun_out/r2ad_bench.err
