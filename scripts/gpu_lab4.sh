#!/bin/bash
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 -i 0 > /tmp/smi.log &
SMI=$!
sleep 1
timeout 300 python scripts/e2e_diag.py 2>&1 | grep -v Warn | tail -5
kill $SMI
wc -l /tmp/smi.log
