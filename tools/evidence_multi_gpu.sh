#!/bin/bash
# multi-GPU evidence: the driver's launch line at N = all GPUs of the box (default workload + c5 + reference arm), then N/2, then the concurrent PCIe probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
run() { # n, tag, extra args
  n=$1; tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/ev_${tag}_n$n.json 2> gpurun_out/ev_${tag}_n$n.err; echo "bench $tag N=$n rc=$?"
}
run $N default --steps 20 --warmup 3 --no-extra
run $N c5 --workload c5 --steps 20 --warmup 3 --no-extra --no-cpu
run $N ref --impl reference --steps 3 --warmup 1
if [ $N -ge 4 ]; then
  run $((N/2)) default --steps 20 --warmup 3 --no-extra --no-cpu
  run $((N/2)) c5 --workload c5 --steps 20 --warmup 3 --no-extra --no-cpu
fi
timeout 600 python tools/pcie_probe_concurrent.py --seconds 1.0 > gpurun_out/ev_pcie_concurrent.json 2> gpurun_out/ev_pcie_concurrent.err; echo "pcie probe rc=$?"
cat gpurun_out/ev_pcie_concurrent.err | tail -6
nvidia-smi topo -m > gpurun_out/ev_topo.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ev_*_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.1f'%d['value'], 'n', d['n_gpus'], d.get('scaling'), 'frac', round((d.get('roofline') or {}).get('frac',0),3), 'e2e', round((d.get('e2e') or {}).get('value',0),1))
    except Exception as e: print(f, 'ERR', e, open(f).read()[:300])
PY
