#!/bin/bash
# dram bytes of the walker vs resident chains / prefetch / hints
for cfg in "32768 0 0" "24576 0 0" "16384 0 0" "16384 0 1" "16384 8 0" "12288 0 0" "12288 0 1" "8192 0 0" "8192 0 1"; do
  set -- $cfg
  KVM_WALK_PREFETCH=$2 KVM_WALK_HINTS=$3 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cnsm_walk -s 1 -c 1 --csv python tools/one_query.py 1e8 1024 $1 5.0 2 2>/dev/null | grep -E "cnsm_walk" | awk -F'","' -v c="$cfg" '{printf "%s | %s %s\n", c, $(NF-2), $NF}' | tr '\n' ';'
  echo
done
