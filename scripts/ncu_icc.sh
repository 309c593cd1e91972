#!/bin/bash
# instruction-cache view of the eval kernel for each variant line of $1
M=gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__average_cache_request_hit_rate.pct,gcc__xbar2gcc_sectors.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
export CB_MAX_ROUNDS=64
while read -r line; do
  [ -z "$line" ] && continue
  echo "== $line"
  env $line ncu --metrics $M --clock-control none -k regex:k_eval_bsimcmg107_nmos -s 40 -c 1 --csv python scripts/first_perf.py 16384 adaptive 6e-8 2>&1 | grep '^"' | awk -F'","' 'NR>1{print $(NF-2), $NF}' | tr -d '"'
done < "$1"
