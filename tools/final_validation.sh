#!/bin/bash
# Round-end style validation on an N-GPU box: GPU tests, smoke, the reference arm and the bench at 1/2/4/8 GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err; tail -c 400 gpurun_out/final_ref_n1.json; echo
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "N=1 rc=$?"; tail -c 300 gpurun_out/final_bench_n1.err
for n in 2 4 8; do
  if [ "$n" -le "$NG" ]; then
    timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/final_bench_n$n.json 2> gpurun_out/final_bench_n$n.err; echo "N=$n rc=$?"
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, "value %.4e ms %.3f e2e %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), "frac", round(d["roofline"]["frac"], 3), "whole", round(d["roofline"]["whole_msm_frac"], 3))
    if d.get("strong"):
        print("   strong", {k: (round(v.get("ms_sharded", 0), 3), round(v.get("speedup_vs_n1", 0), 2)) for k, v in d["strong"].items()})
    if d.get("strong_in_library"):
        print("   in-library", {k: {kk: round(vv, 2) for kk, vv in v.items() if isinstance(vv, float)} for k, v in d["strong_in_library"].items() if isinstance(v, dict)})
PY
