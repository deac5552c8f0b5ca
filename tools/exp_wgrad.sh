export ENS_BWD_TC_SPLIT=1
for e in ${EXPS:-0 1 2 3 4 5}; do
  echo "== ENS_WGRAD_EXP=$e"
  ENS_WGRAD_EXP=$e timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad_tc" --csv python tools/prof_step.py 2>/dev/null | grep wgrad_tc | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
  echo
done
