#!/bin/bash
# compute-sanitizer passes over the new round-2 kernels at sizes that finish in seconds
# (run on the GPU box; logs land in gpurun_out/ and are summarised into profiles/).
#   bash profiles/run_sanitizer.sh [1|2]     # 2: also the peer-memory protocol on two GPUs
set -u
out=gpurun_out
mkdir -p $out
SMALL='tests/test_gpu_amg.py -k dist_amg_single_rank and 100'
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 \
    python -m pytest tests/test_gpu_amg.py -q -x -k "dist_amg_single_rank and 100" > $out/r2_sanitizer_${tool}_amg.log 2>&1
  tail -4 $out/r2_sanitizer_${tool}_amg.log | cut -c1-200
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 \
    python -m pytest tests/test_gpu_assembly.py -q -x -k "doc_netlists or corner or grid" > $out/r2_sanitizer_${tool}_assembly.log 2>&1
  tail -4 $out/r2_sanitizer_${tool}_assembly.log | cut -c1-200
done
if [ "${1:-1}" = "2" ]; then
  for tool in racecheck synccheck; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      --no-python compute-sanitizer --tool $tool --print-limit 5 python tests/dist_check.py 100 > $out/r2_sanitizer_${tool}_dist2.log 2>&1
    grep -E "DIST_CHECK|ERROR SUMMARY|RACECHECK SUMMARY" $out/r2_sanitizer_${tool}_dist2.log | head -6
  done
fi
