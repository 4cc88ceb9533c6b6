#!/bin/bash
# env.step time of the q-learning loop for several phase schedules (PKB_ENV_PHASE_ROUNDS)
for r in "6,12,32,0" "4,8,16,32,0" "3,6,12,24,0" "3,5,8,16,32,0" "2,4,8,16,32,0" "4,6,10,16,32,64,0" "8,0" "4,0" "0"; do
  echo -n "$r: "; PKB_ENV_PHASE_ROUNDS=$r python profiles/ql_breakdown.py 2>&1 | grep -E "env.step|^total" | awk '{printf "%s %s ms | ", $1, ($1=="total"?$2:$7)}'; echo
done
