#!/bin/bash
# final N=1 evidence of the round on one build: tests, sanitizers, bench lines of every workload + the reference arm, probes, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ev_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/ev_pytest.log
tail -3 gpurun_out/ev_pytest.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_bench_reference.json 2> gpurun_out/ev_bench_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/ev_bench_default.json 2> gpurun_out/ev_bench_default.err; echo "bench default rc=$?"
for w in c3 c4 c5; do
  timeout 900 python bench.py --workload $w --no-cpu > gpurun_out/ev_bench_$w.json 2> gpurun_out/ev_bench_$w.err; echo "bench $w rc=$?"
done
timeout 600 python tools/varint_probe.py > gpurun_out/ev_varint_probe.json 2> gpurun_out/ev_varint_probe.err; echo "varint probe rc=$?"
B200TFS_FUSED_VARINT=1 timeout 600 python tools/varint_probe.py > gpurun_out/ev_varint_probe_fused.json 2> gpurun_out/ev_varint_probe_fused.err; echo "varint probe (single-pass experiment) rc=$?"
timeout 300 python tools/pcie_probe.py > gpurun_out/ev_pcie_probe.log 2>&1
timeout 600 python tools/decode_latency_probe.py > gpurun_out/ev_decode_latency.json 2> gpurun_out/ev_decode_latency.err; echo "decode latency probe rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --deselect tests/test_device_api_gpu.py::test_one_gib_tensor --deselect tests/test_golden_gpu.py::test_single_pass_varint_kernels_agree > gpurun_out/ev_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/ev_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_golden_gpu.py tests/test_device_api_gpu.py tests/test_host_pipeline_gpu.py -m gpu -x -q -k "varint or round_trip or alignment or staged or padding or deferred or template or sliced or narrow" > gpurun_out/ev_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/ev_racecheck.log
bash tools/evidence_ncu.sh > gpurun_out/ev_prof.log 2>&1; tail -5 gpurun_out/ev_prof.log
python - <<'PY'
import json
for f in ('default','c3','c4','c5','reference'):
    try:
        d=json.loads(open('gpurun_out/ev_bench_%s.json'%f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}; e=d.get('e2e') or {}
        print(f, 'value',round(d['value'],2), 'frac', round(r.get('frac',0),3), 'enc', round((r.get('encode') or {}).get('frac',0),3), 'dec', round((r.get('decode') or {}).get('frac',0),3), d.get('mode') or (d.get('config') or {}).get('timed_region'), 'e2e', round(e.get('value',0),2), 'one', (e.get('one_pair_at_a_time') or {}).get('us_per_pair'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
