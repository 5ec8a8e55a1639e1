#!/bin/bash
# retry until a box is free
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout 600 -- 'mkdir -p gpurun_out; (timeout 200 python -m pytest tests/test_wide_gpu.py -x -q -k "512-200" 2>&1 | tail -40) > gpurun_out/wide_t0.log; (RSSM_WIDE_DESC_SWAP=1 timeout 200 python -m pytest tests/test_wide_gpu.py -x -q -k "512-200" 2>&1 | tail -40) > gpurun_out/wide_t1.log; (timeout 120 python profiles/src/wide_bench.py 2>&1 | tail -5) > gpurun_out/wide_b0.log; echo done' > gpurun_out/wide_call0.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/wide_call0.log; then break; fi
  sleep 120
done
echo finished rc=$rc
