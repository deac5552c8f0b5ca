#!/bin/bash
# retry until the pod has a slot
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout 1500 -- 'timeout 600 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo rc=$?; timeout 400 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo rc=$?' > gpurun_out/final_n1.call.log 2>&1
  if grep -q "status=ok" gpurun_out/final_n1.call.log; then echo ok; exit 0; fi
  sleep 150
done
echo gave_up
