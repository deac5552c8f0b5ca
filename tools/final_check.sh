#!/bin/bash
for i in $(seq 1 10); do
  /usr/local/graft/bin/gpurun --timeout 900 -- 'timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; timeout 200 python tools/prof_track_host.py 2>&1 | grep "wall ms"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1' > gpurun_out/final_check.log 2>&1
  if grep -q "status=ok" gpurun_out/final_check.log; then echo ok; exit 0; fi
  sleep 120
done
echo gave_up
